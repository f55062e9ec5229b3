// Micro-benchmark 4: how fast can a persistent kernel STREAM a row-major bf16 matrix [M][K] through shared memory with
// TMA boxes of (64 columns = 128 B) x ROWS rows, the access pattern of the skinny LoRA GEMMs?  One producer thread per
// CTA, STAGES-deep ring, a consumer warp that only waits and releases.  Sweeps box rows / stages / CTAs per SM.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// work item = (row tile, k-block); mode 0: a CTA owns row tiles and walks K (like lora_down); mode 1: work items are
// handed out so that consecutive CTAs take consecutive k-blocks of the same row tile (wide contiguous rows in flight)
__global__ void __launch_bounds__(64) stream_kernel(const __grid_constant__ CUtensorMap map, int M, int K, int rows, int stages,
                                                    int mode, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t stage_bytes = rows * 128;
    const uint32_t bars = base + stages * stage_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (stages + s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int m_tiles = (M + rows - 1) / rows, kbs = K / 64;
    // work enumeration (32-bit): every CTA walks the same kind of sequence in both roles
    //   mode 0: own row tiles (blockIdx, +grid, ...), all k-blocks in order
    //   mode 1: items c, c + grid, ... of the flat (row tile, k-block) list
    //   mode S >= 2: group g = blockIdx / S owns row tiles g, g + G, ...; member j takes k-blocks j, j + S, ...
    const int S = mode >= 2 ? mode : 1, G = gridDim.x / S, g = blockIdx.x / S, j = blockIdx.x % S;
    auto walk = [&](auto&& body) {
        if (mode == 1) {
            int mt = blockIdx.x / kbs, kb = blockIdx.x % kbs;
            const int dmt = gridDim.x / kbs, dkb = gridDim.x % kbs;
            while (mt < m_tiles) {
                body(mt, kb);
                mt += dmt; kb += dkb;
                if (kb >= kbs) { kb -= kbs; ++mt; }
            }
        } else if (mode == 0) {
            for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x)
                for (int kb = 0; kb < kbs; ++kb) body(mt, kb);
        } else {
            if (g >= G) return;
            for (int mt = g; mt < m_tiles; mt += G)
                for (int kb = j; kb < kbs; kb += S) body(mt, kb);
        }
    };
    if (threadIdx.x == 0) {   // producer
        int s = 0; uint32_t ph = 0;
        walk([&](int mt, int kb) {
            mbar_wait(bars + 8 * (stages + s), ph ^ 1u);
            mbar_expect(bars + 8 * s, stage_bytes);
            tma_load_2d(base + s * stage_bytes, &map, bars + 8 * s, kb * 64, mt * rows);
            if (++s == stages) { s = 0; ph ^= 1u; }
        });
    } else if (threadIdx.x == 32) {   // consumer: wait + touch one word + release
        int s = 0; uint32_t ph = 0; unsigned long long acc = 0;
        walk([&](int, int) {
            mbar_wait(bars + 8 * s, ph);
            acc += *reinterpret_cast<volatile uint32_t*>(smem + (base - smem_u32(smem)) + s * stage_bytes);
            mbar_arrive(bars + 8 * (stages + s));
            if (++s == stages) { s = 0; ph ^= 1u; }
        });
        sink[blockIdx.x] = acc;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fnp;
    const int M = 16384;
    unsigned long long* sink; cudaMalloc(&sink, 4096 * 8);
    for (int K : {4096, 14336}) {
        void* x; cudaMalloc(&x, (size_t)M * K * 2); cudaMemset(x, 1, (size_t)M * K * 2);
        for (int rows : {128}) for (int stages : {6}) for (int cps : {1}) for (int mode : {0, 1, 2, 4, 8, 16, 37, 74}) {
            const size_t smem = 1024 + (size_t)stages * rows * 128 + 16 * stages + 64;
            if (smem * cps > 227 * 1024) continue;
            CUtensorMap map;
            cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
            if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
            cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const int grid = 148 * cps;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                stream_kernel<<<grid, 64, smem>>>(map, M, K, rows, stages, mode, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
            }
            cudaError_t err = cudaGetLastError();
            printf("K %5d rows %3d stages %2d ctas/sm %d mode %d: %7.1f us  %6.0f GB/s  (%s)\n", K, rows, stages, cps, mode, best * 1e3,
                   (double)M * K * 2 / best / 1e6, cudaGetErrorString(err));
        }
        cudaFree(x);
    }
    return 0;
}
