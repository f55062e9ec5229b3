// Micro-benchmark: TMEM -> register read throughput per SM (tcgen05.ld 32x32b), 4 warps = 128 lanes x 512 columns.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o causal-unified-language-vision_b200/build/tmem_ld_bench tests/micro/tmem_ld_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldwait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: x32, wait after each.  MODE 1: two x32 in flight, one wait.  MODE 2: x16, wait after each.
// NW warps active (1, 2 or 4).
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, uint32_t* sink, int reps, int nw) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nw) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (MODE == 0) {
#pragma unroll 1
                for (int c = 0; c < 16; ++c) {
                    uint32_t v[32];
                    ld32(base + c * 32, v);
                    ldwait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc ^= v[j];
                }
            } else if (MODE == 1) {
#pragma unroll 1
                for (int c = 0; c < 16; c += 2) {
                    uint32_t v[32], w[32];
                    ld32(base + c * 32, v);
                    ld32(base + c * 32 + 32, w);
                    ldwait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc ^= v[j] ^ w[j];
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < 32; ++c) {
                    uint32_t v[16];
                    ld16(base + c * 16, v);
                    ldwait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc ^= v[j];
                }
            }
        }
        t1 = clock64();
    }
    __syncthreads();
    if (threadIdx.x % 32 == 0 && warp < nw) out[blockIdx.x * 4 + warp] = t1 - t0;
    sink[blockIdx.x * 128 + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 148 * 4 * sizeof(long long));
    cudaMalloc(&sink, 148 * 128 * 4);
    const int reps = 200;
    for (int mode = 0; mode < 3; ++mode)
        for (int nw = 1; nw <= 4; nw *= 2) {
            cudaMemset(out, 0, 148 * 4 * sizeof(long long));
            if (mode == 0) k<0><<<148, 128>>>(out, sink, reps, nw);
            if (mode == 1) k<1><<<148, 128>>>(out, sink, reps, nw);
            if (mode == 2) k<2><<<148, 128>>>(out, sink, reps, nw);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[4];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            // bytes read by the SM per rep: nw warps x 32 lanes x 512 cols x 4 B
            double cyc = (double)h[0] / reps;
            printf("mode %d warps %d: %.0f cycles per 512-column sweep per warp, SM read rate %.1f B/clk (%s)\n", mode, nw, cyc,
                   nw * 32.0 * 512 * 4 / cyc, cudaGetErrorString(e));
        }
    return 0;
}
