// Micro-benchmark 3: the V6 epilogue loop (TMEM 32x64 group -> bf16 -> swizzled smem tile -> read back -> global) in
// isolation, 4 warps per CTA, one CTA per SM; stages switched on one at a time.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
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
__device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(128, 1) k(long long* out, __nv_bfloat16* D, long long ldd, int reps, int mode) {
    __shared__ uint32_t slot;
    __shared__ __align__(1024) uint8_t stage[4 * 4096];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t_warp = slot + ((uint32_t)(warp * 32) << 16);
    const uint32_t stg = (uint32_t)__cvta_generic_to_shared(stage) + warp * 4096;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const int m0 = (blockIdx.x * reps + r) % 64 * 128;   // a different 128-row output block per rep
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            uint32_t o[32];
            {
                uint32_t v[32], w[32];
                ld32(t_warp + q * 64, v);
                ld32(t_warp + q * 64 + 32, w);
                ldwait();
                if (mode >= 1) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        o[j] = pack(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                        o[16 + j] = pack(__uint_as_float(w[2 * j]), __uint_as_float(w[2 * j + 1]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] = v[j] ^ w[j];
                }
            }
            if (mode < 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= o[j];
                continue;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg + lane * 128 + ((uint32_t)(j ^ (lane & 7)) << 4)),
                             "r"(o[4 * j]), "r"(o[4 * j + 1]), "r"(o[4 * j + 2]), "r"(o[4 * j + 3]) : "memory");
            __syncwarp();
            if (mode < 3) continue;
            const int ch = lane & 7;
            uint4 val[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + (lane >> 3);
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(val[i].x), "=r"(val[i].y), "=r"(val[i].z), "=r"(val[i].w)
                             : "r"(stg + rr * 128 + ((uint32_t)(ch ^ (rr & 7)) << 4)));
            }
            if (mode < 4) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc ^= val[i].x ^ val[i].y ^ val[i].z ^ val[i].w;
                continue;
            }
            __nv_bfloat16* gp = D + (long long)(m0 + warp * 32 + (lane >> 3)) * ldd + (blockIdx.x % 8) * 512 + q * 64 + ch * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(gp + (long long)i * 4 * ldd) = val[i];
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345u) D[0] = __float2bfloat16(1.f);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
    long long* out; __nv_bfloat16* D;
    const long long ldd = 4096;
    cudaMalloc(&out, 148 * sizeof(long long));
    cudaMalloc(&D, 8192 * ldd * 2);
    const int reps = 50;
    for (int mode = 0; mode <= 4; ++mode) {
        k<<<148, 128>>>(out, D, ldd, reps, mode);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("mode %d: %.0f cycles per 128x512 accumulator (cta0), max %.0f (%s)\n", mode, (double)h[0] / reps, mx / reps, cudaGetErrorString(e));
    }
    return 0;
}
