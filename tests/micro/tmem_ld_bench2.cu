// Micro-benchmark 2: TMEM read sweep (16 x tcgen05.ld 32x32b.x32 + wait) by 4 warps while other conditions of the
// real GEMM kernel are emulated one at a time:
//   spin=1   8 extra warps poll an mbarrier that is only completed at the end (like stalled decode warps)
//   pair=1   launched as a 2-CTA cluster with a cta_group::2 TMEM allocation
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
__device__ __forceinline__ void ldwait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}

template <int PAIR>
__global__ void __launch_bounds__(384, 1) k(long long* out, uint32_t* sink, int reps, int spin, int first_warp) {
    __shared__ uint32_t slot;
    __shared__ __align__(8) unsigned long long bar;
    const int warp = threadIdx.x >> 5;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(4));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp >= first_warp && warp < first_warp + 4) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll 1
            for (int c = 0; c < 16; ++c) {
                uint32_t v[32];
                ld32(base + c * 32, v);
                ldwait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= v[j];
            }
        }
        t1 = clock64();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
    } else if (spin) {
        while (!try_wait(bar_a, 0)) {}
    }
    __syncthreads();
    if (threadIdx.x % 32 == 0 && warp == first_warp) out[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 384 + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
    if (warp == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
    }
}

int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 148 * sizeof(long long));
    cudaMalloc(&sink, 148 * 384 * 4);
    const int reps = 100;
    for (int pair = 0; pair < 2; ++pair)
        for (int spin = 0; spin < 2; ++spin)
            for (int fw = 0; fw <= 4; fw += 4) {
                cudaMemset(out, 0, 148 * sizeof(long long));
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(148); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaError_t e;
                if (pair) e = cudaLaunchKernelEx(&cfg, k<1>, out, sink, reps, spin, fw);
                else e = cudaLaunchKernelEx(&cfg, k<0>, out, sink, reps, spin, fw);
                cudaError_t e2 = cudaDeviceSynchronize();
                long long h[2];
                cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                printf("pair %d spin %d warps %d-%d: %.0f cycles per 16-chunk sweep (%s / %s)\n", pair, spin, fw, fw + 3,
                       (double)h[0] / reps, cudaGetErrorString(e), cudaGetErrorString(e2));
            }
    return 0;
}
