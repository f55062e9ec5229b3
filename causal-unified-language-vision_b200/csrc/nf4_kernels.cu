// Stand-alone NF4 kernels for sm_100a: blockwise decode (bit-exactness probe of the
// in-register decode used by the GEMM main loops), blockwise quantise (+ double quant of
// the absmax), counter-based dropout mask.  HBM-bound SIMT kernels: 128-bit / fully
// coalesced accesses, grids sized in multiples of the SM count.
//
// Replaces (reference side, third-party):
//   cdequantize_blockwise_bf16_nf4 (+ nested cdequantize_blockwise_fp32)  -> b2q_nf4_decode
//   cquantize_blockwise_*_nf4 (+ cquantize_blockwise_fp32 for the absmax) -> b2q_nf4_quantize
#include <cstdlib>

#include "b2q_decode.cuh"
#include "b2q_internal.h"

namespace b2q {

// ------------------------------------------------------------------ decode ----
// Each lane decodes one packed 32-bit word (8 weights) per step; consecutive lanes take
// consecutive words, so loads (128 B / warp) and stores (512 B / warp) are coalesced.
template <int ALGO>
__global__ void __launch_bounds__(256) nf4_decode_kernel(const uint32_t* __restrict__ packed, AbsmaxSrc am,
                                                         const float* __restrict__ code16_g,
                                                         uint4* __restrict__ out, long long n_words) {
    float code16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) code16[i] = __ldg(code16_g + i);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long wi = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; wi < n_words; wi += stride) {
        const uint32_t w = __ldg(packed + wi);
        const float a = load_absmax(am, wi >> 3);  // 8 words = 64 weights = one block
        uint32_t o[4];
        if constexpr (ALGO == 1) {
            Nf4Lut lut;
            nf4_build_lut(code16, a, lut);
            nf4_decode_word(w, lut, o);
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t byte = (w >> (8 * b)) & 0xFFu;
                float lo = 0.f, hi = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) {  // register-resident table, select by compare
                    lo = ((byte >> 4) == c) ? code16[c] : lo;
                    hi = ((byte & 15u) == c) ? code16[c] : hi;
                }
                o[b] = pack_bf16x2(__fmul_rn(lo, a), __fmul_rn(hi, a));
            }
        }
        out[wi] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}


// -------------------------------------------------------------------- GEMV ----
// y[m, n] = sum_k x[m, k] * dequant(W)[n, k] for a handful of token rows (single-token decoding in `generate`):
// HBM-bound on the packed weight (N*K/2 bytes), so no tensor cores -- one warp per output row, each lane streams
// 16 packed bytes (half a quantisation block, one absmax) per step with 128-bit loads, decodes them with the same
// pre-scaled LUT as the GEMM main loops (bit-identical bf16 weights), accumulates in fp32 and the warp reduces with
// shuffles.  Replaces bitsandbytes `cgemm_4bit_inference_naive_bf16` (kgemm_4bit_inference_naive).
template <int MROWS, int RPW, int U, int MINB>
__global__ void __launch_bounds__(128, MINB) nf4_gemv_kernel(const __nv_bfloat16* __restrict__ x, const uint4* __restrict__ packed,
                                                       AbsmaxSrc am, const float* __restrict__ code16_g,
                                                       __nv_bfloat16* __restrict__ y, int M, int N, int K) {
    // RPW: output rows per warp (every x chunk fetched from L1 is used for all of them); U: chunks in flight per lane
    // and row (memory-level parallelism); MINB: resident blocks per SM the register allocation must allow
    float code16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) code16[i] = __ldg(code16_g + i);
    const int lane = threadIdx.x & 31;
    const int n0 = (blockIdx.x * 4 + (threadIdx.x >> 5)) * RPW;
    if (n0 >= N) return;
    const int chunks = K / 32;                       // 16-byte chunks (32 weights) per row
    float acc[RPW][MROWS];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int m = 0; m < MROWS; ++m) acc[r][m] = 0.f;
    // software pipeline: the packed bytes and absmax of batch i+1 are in flight while batch i is decoded
    uint4 q[RPW][U], qn[RPW][U];
    float a[RPW][U], an[RPW][U];
    auto fetch = [&](int c0, uint4 (&qq)[RPW][U], float (&aa)[RPW][U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + 32 * u;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const bool ok = c < chunks && n0 + r < N;
                qq[r][u] = ok ? __ldg(packed + static_cast<long long>(n0 + r) * chunks + c) : make_uint4(0u, 0u, 0u, 0u);
                aa[r][u] = ok ? load_absmax(am, static_cast<long long>(n0 + r) * (K / 64) + (c >> 1)) : 0.f;
            }
        }
    };
    fetch(lane, q, a);
    for (int c0 = lane; c0 < chunks; c0 += 32 * U) {
        fetch(c0 + 32 * U, qn, an);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + 32 * u;
            if (c >= chunks) break;
            uint32_t d[RPW][16];                     // 32 bf16 weights per row, element order
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                Nf4Lut lut;
                nf4_build_lut(code16, a[r][u], lut);
                const uint32_t w[4] = {q[r][u].x, q[r][u].y, q[r][u].z, q[r][u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t o[4];
                    nf4_decode_word(w[j], lut, o);
                    d[r][4 * j] = o[0]; d[r][4 * j + 1] = o[1]; d[r][4 * j + 2] = o[2]; d[r][4 * j + 3] = o[3];
                }
            }
#pragma unroll
            for (int m = 0; m < MROWS; ++m) {
                if (m < M) {
                    const uint4* xp = reinterpret_cast<const uint4*>(x + static_cast<long long>(m) * K + c * 32);
                    float s[RPW];
#pragma unroll
                    for (int r = 0; r < RPW; ++r) s[r] = 0.f;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint4 xv = __ldg(xp + t);
                        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float xl = __uint_as_float(xw[e] << 16), xh = __uint_as_float(xw[e] & 0xFFFF0000u);
#pragma unroll
                            for (int r = 0; r < RPW; ++r) {
                                const uint32_t wv = d[r][4 * t + e];
                                s[r] = fmaf(__uint_as_float(wv << 16), xl, s[r]);
                                s[r] = fmaf(__uint_as_float(wv & 0xFFFF0000u), xh, s[r]);
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RPW; ++r) acc[r][m] += s[r];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int r = 0; r < RPW; ++r) { q[r][u] = qn[r][u]; a[r][u] = an[r][u]; }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int m = 0; m < MROWS; ++m) {
            float v = acc[r][m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && m < M && n0 + r < N) y[static_cast<long long>(m) * N + n0 + r] = __float2bfloat16_rn(v);
        }
}

// Tensor-core formulation of the same GEMV (default since round 2: 35-37 us for every M <= 8 at the 14336 x 4096 shapes
// against 34 / 58 / 99 us of the SIMT kernel at M = 1 / 4 / 8; B2Q_GEMV_CFG=1 selects the SIMT kernel).  The SIMT kernel above
// spends one FMA and one unpack per weight and token on top of the decode; here the decoded bf16x2 registers ARE the A
// fragments of mma.sync.m16n8k16 (16 weight rows x 16 k), the tokens are the 8 columns of the B fragment, and the
// accumulation costs no ALU work at any M <= 8.  A block owns 16 weight rows; its 8 warps take 128-weight k-steps in
// turn (all of a warp's packed loads are in flight at once) and their fp32 partial sums are added in a fixed order
// through shared memory (deterministic).  Lane (g = lane/4, t = lane%4) decodes the t-th 32-weight chunk of rows g and
// g+8 in each k-step; MMA j of the step contracts weights 4j..4j+3 of every lane's chunk, i.e. the MMA's k index
// {2t, 2t+1, 2t+8, 2t+9} stands for elements {4j, 4j+1, 4j+2, 4j+3} of chunk t -- the B fragment is loaded from x with
// the same permutation, so the sum is over the true k.  Decoded weights are bit-identical to b2q_nf4_decode.
__device__ __forceinline__ void mma_bf16_m16n8k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                  uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int U>
__global__ void __launch_bounds__(256, 2) nf4_gemv_mma_kernel(const __nv_bfloat16* __restrict__ x,
                                                              const uint4* __restrict__ packed, AbsmaxSrc am,
                                                              const float* __restrict__ code16_g,
                                                              __nv_bfloat16* __restrict__ y, int M, int N, int K) {
    __shared__ float red[8][128];
    float code16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) code16[i] = __ldg(code16_g + i);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.x * 16;
    const int chunks = K / 32;              // 16-byte chunks (32 weights) per row
    const int steps = (chunks + 3) / 4;     // k-steps of 4 chunks (one per t)
    const int row_a = n0 + g, row_b = n0 + g + 8;
    const bool ok_a = row_a < N, ok_b = row_b < N;
    const bool tok = g < M;                 // column g of the B fragment is token g
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s0 = warp; s0 < steps; s0 += 8 * U) {
        uint4 qa[U], qb[U];
        float aa[U], ab[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = 4 * (s0 + 8 * u) + t;
            const bool ok = c < chunks;     // also false for k-steps past the end
            qa[u] = (ok && ok_a) ? __ldg(packed + static_cast<long long>(row_a) * chunks + c) : zero4;
            qb[u] = (ok && ok_b) ? __ldg(packed + static_cast<long long>(row_b) * chunks + c) : zero4;
            aa[u] = (ok && ok_a) ? load_absmax(am, static_cast<long long>(row_a) * (K / 64) + (c >> 1)) : 0.f;
            ab[u] = (ok && ok_b) ? load_absmax(am, static_cast<long long>(row_b) * (K / 64) + (c >> 1)) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (s0 + 8 * u >= steps) break;             // warp-uniform
            const int c = 4 * (s0 + 8 * u) + t;
            uint32_t xr[16];                            // x[token g][c*32 .. c*32+31] as bf16x2 words
            if (tok && c < chunks) {
                const uint4* xp = reinterpret_cast<const uint4*>(x + static_cast<long long>(g) * K + c * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 v = __ldg(xp + i);
                    xr[4 * i] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) xr[i] = 0u;
            }
            uint32_t da[16], db[16];                    // 32 decoded weights of rows g and g+8, element order
            {
                Nf4Lut lut;
                nf4_build_lut(code16, aa[u], lut);
                const uint32_t w[4] = {qa[u].x, qa[u].y, qa[u].z, qa[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t o[4];
                    nf4_decode_word(w[j], lut, o);
                    da[4 * j] = o[0]; da[4 * j + 1] = o[1]; da[4 * j + 2] = o[2]; da[4 * j + 3] = o[3];
                }
            }
            {
                Nf4Lut lut;
                nf4_build_lut(code16, ab[u], lut);
                const uint32_t w[4] = {qb[u].x, qb[u].y, qb[u].z, qb[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t o[4];
                    nf4_decode_word(w[j], lut, o);
                    db[4 * j] = o[0]; db[4 * j + 1] = o[1]; db[4 * j + 2] = o[2]; db[4 * j + 3] = o[3];
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                mma_bf16_m16n8k16(acc, da[2 * j], db[2 * j], da[2 * j + 1], db[2 * j + 1], xr[2 * j], xr[2 * j + 1]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) red[warp][lane * 4 + i] = acc[i];
    __syncthreads();
    if (threadIdx.x < 128) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        // accumulator element i of lane (gg, tt): row gg + 8 * (i >> 1), token 2 * tt + (i & 1)
        const int ln = threadIdx.x >> 2, i = threadIdx.x & 3;
        const int row = n0 + (ln >> 2) + 8 * (i >> 1), token = 2 * (ln & 3) + (i & 1);
        if (token < M && row < N) y[static_cast<long long>(token) * N + row] = __float2bfloat16_rn(v);
    }
}

// ---------------------------------------------------------------- quantise ----
__device__ __forceinline__ uint32_t nf4_code_of(float x) {
    // number of thresholds strictly below x (== bitsandbytes' dQuantizeNF4 comparison tree); NaN -> 0
    const float t[15] = {-0.8480964004993439f,  -0.6106329262256622f,  -0.4599952697753906f, -0.33967943489551544f,
                         -0.23460740596055984f, -0.13791173323988914f, -0.045525018125772476f, 0.03979014977812767f,
                         0.1202552504837513f,   0.2035212516784668f,   0.2920137718319893f,  0.3893125355243683f,
                         0.5016634166240692f,   0.6427869200706482f,   0.8614784181118011f};
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 15; ++i) c += (x > t[i]) ? 1u : 0u;
    return c;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// One warp per 64-element block: lane l owns elements 2l, 2l+1 -> byte l of the block.
template <typename T>
__global__ void __launch_bounds__(256) nf4_quantize_kernel(const T* __restrict__ w, uint8_t* __restrict__ packed,
                                                           float* __restrict__ absmax, long long n) {
    const int lane = threadIdx.x & 31;
    const long long nblocks = (n + 63) >> 6;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long blk = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; blk < nblocks;
         blk += warps) {
        const long long i0 = blk * 64 + 2 * lane;
        const float v0 = (i0 < n) ? to_f32<T>(w[i0]) : 0.f;
        const float v1 = (i0 + 1 < n) ? to_f32<T>(w[i0 + 1]) : 0.f;
        float a = fmaxf(fabsf(v0), fabsf(v1));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
        const float inv = __fdiv_rn(1.0f, a);  // inf for an all-zero block -> 0*inf = NaN -> code 0
        const uint32_t c0 = nf4_code_of(__fmul_rn(v0, inv));
        const uint32_t c1 = nf4_code_of(__fmul_rn(v1, inv));
        if (i0 < n) packed[i0 >> 1] = static_cast<uint8_t>((c0 << 4) | c1);
        if (lane == 0) absmax[blk] = a;
    }
}

// Deterministic mean of the absmax vector (fp64 accumulate, fixed tree), single CTA.
__global__ void __launch_bounds__(1024) absmax_mean_kernel(const float* __restrict__ absmax, long long nblocks,
                                                           float* __restrict__ offset_out) {
    __shared__ double sh[1024];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < nblocks; i += 1024) acc += static_cast<double>(absmax[i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *offset_out = static_cast<float>(sh[0] / static_cast<double>(nblocks));
}

// 8-bit blockwise(256) quantisation of (absmax - offset) to the nearest entry of the
// sorted 256-entry map (ties -> lower index).  One CTA of 256 threads per block.
__global__ void __launch_bounds__(256) absmax_quantize_kernel(const float* __restrict__ absmax,
                                                              const float* __restrict__ offset_p,
                                                              const float* __restrict__ code256,
                                                              uint8_t* __restrict__ q, float* __restrict__ absmax2,
                                                              long long nblocks) {
    __shared__ float code[256];
    __shared__ float red[8];
    code[threadIdx.x] = code256[threadIdx.x];
    const float offset = *offset_p;
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    const float v = (i < nblocks) ? __fsub_rn(absmax[i], offset) : 0.f;
    float a = fabsf(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    a = red[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) a = fmaxf(a, red[k]);
    if (threadIdx.x == 0) absmax2[blockIdx.x] = a;
    float x = __fmul_rn(v, __fdiv_rn(1.0f, a));
    if (x != x) x = 0.f;
    // first index with code[idx] >= x, clamped to [1,255]
    int lo = 0, hi = 256;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (code[mid] < x) lo = mid + 1; else hi = mid;
    }
    int up = min(max(lo, 1), 255);
    int dn = up - 1;
    const int pick = ((code[up] - x) < (x - code[dn])) ? up : dn;
    if (i < nblocks) q[i] = static_cast<uint8_t>(pick);
}

// ----------------------------------------------------------------- dropout ----
// keep(m,k) = 15-bit field of hash64(seed, (m*K + k) / 4) >= round(p * 2^15)   (b2q_internal.h dropout_keep; counter
// based, so forward, backward and the oracle -- oracle/dropout.py -- all regenerate the same mask from (seed, p)).
__global__ void __launch_bounds__(256) dropout_mask_kernel(uint8_t* __restrict__ mask, long long n,
                                                           unsigned long long seed, uint32_t thresh) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        mask[i] = dropout_keep(seed, static_cast<unsigned long long>(i), thresh) ? 1 : 0;
}

// xd = bf16(x * keep / (1 - p)), 8 elements (16 B) per thread-step.
__global__ void __launch_bounds__(256) dropout_apply_kernel(const uint4* __restrict__ x, uint4* __restrict__ xd,
                                                            long long n_vec, unsigned long long seed,
                                                            uint32_t thresh, float scale) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        uint4 v = __ldg(x + i);
        uint32_t* p = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&p[j]);
            const unsigned long long e = static_cast<unsigned long long>(i) * 8 + 2 * j;
            const float a = dropout_keep(seed, e, thresh) ? __bfloat162float(h.x) * scale : 0.f;
            const float b = dropout_keep(seed, e + 1, thresh) ? __bfloat162float(h.y) * scale : 0.f;
            p[j] = pack_bf16x2(a, b);
        }
        xd[i] = v;
    }
}

// dx += bf16(dxl * keep / (1 - p)), 8 elements per thread-step.
__global__ void __launch_bounds__(256) dropout_bwd_add_kernel(uint4* __restrict__ dx, const uint4* __restrict__ dxl,
                                                              long long n_vec, unsigned long long seed,
                                                              uint32_t thresh, float scale) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        uint4 a = dx[i];
        uint4 b = __ldg(dxl + i);
        uint32_t* pa = reinterpret_cast<uint32_t*>(&a);
        uint32_t* pb = reinterpret_cast<uint32_t*>(&b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 ha = *reinterpret_cast<__nv_bfloat162*>(&pa[j]);
            __nv_bfloat162 hb = *reinterpret_cast<__nv_bfloat162*>(&pb[j]);
            const unsigned long long e = static_cast<unsigned long long>(i) * 8 + 2 * j;
            // the reference rounds the masked LoRA gradient to bf16 before the add
            const float l0 = dropout_keep(seed, e, thresh) ? __bfloat162float(hb.x) * scale : 0.f;
            const float l1 = dropout_keep(seed, e + 1, thresh) ? __bfloat162float(hb.y) * scale : 0.f;
            const float r0 = __bfloat162float(ha.x) + __bfloat162float(__float2bfloat16_rn(l0));
            const float r1 = __bfloat162float(ha.y) + __bfloat162float(__float2bfloat16_rn(l1));
            pa[j] = pack_bf16x2(r0, r1);
        }
        dx[i] = a;
    }
}

// out_bf16[i] = bf16( (accumulate ? out[i] : 0) + scale * sum_s partial[s][i] )
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int splits,
                                                              long long n, float scale,
                                                              __nv_bfloat16* __restrict__ out, int accumulate) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float acc = 0.f;
        for (int s = 0; s < splits; ++s) acc += partial[static_cast<long long>(s) * n + i];
        acc *= scale;
        if (accumulate) acc += __bfloat162float(out[i]);
        out[i] = __float2bfloat16_rn(acc);
    }
}

// SM count of the calling thread's current device (cached per device)
static int sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cache[dev] = n > 0 ? n : 148;
    }
    return cache[dev];
}

static int grid_for(long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = static_cast<long long>(sm_count()) * 8;  // 8 resident 256-thread CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

}  // namespace b2q

using namespace b2q;

extern "C" int b2q_nf4_decode(const uint8_t* packed, const float* absmax, const uint8_t* absmax_q,
                              const float* absmax2, const float* code256, float offset, const float* code16,
                              void* out_bf16, int64_t n, int blocksize, int algo, cudaStream_t stream) {
    if (blocksize != 64 || n % 64 != 0) return B2Q_ERR_SHAPE;
    if (n == 0) return 0;
    AbsmaxSrc am{absmax, absmax_q, absmax2, code256, offset};
    if (absmax_q == nullptr && absmax == nullptr) return B2Q_ERR_ARG;
    const long long n_words = n / 8;
    const int grid = grid_for(n_words, 256);
    if (algo == 1)
        nf4_decode_kernel<1><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(packed), am, code16,
                                                       reinterpret_cast<uint4*>(out_bf16), n_words);
    else
        nf4_decode_kernel<0><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(packed), am, code16,
                                                       reinterpret_cast<uint4*>(out_bf16), n_words);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}


extern "C" int b2q_gemv_4bit(const void* x_bf16, const b2q_nf4_weight* w, void* y_bf16, int M, int N, int K,
                             cudaStream_t stream) {
    if (M == 0) return 0;
    if (x_bf16 == nullptr || y_bf16 == nullptr || w == nullptr || w->packed == nullptr || w->code16 == nullptr)
        return B2Q_ERR_ARG;
    if (w->absmax_q == nullptr && w->absmax == nullptr) return B2Q_ERR_ARG;
    if (M < 0 || M > 8 || K % 64 != 0 || N <= 0) return B2Q_ERR_SHAPE;
    if (((reinterpret_cast<uintptr_t>(x_bf16) | reinterpret_cast<uintptr_t>(w->packed)) & 15) != 0) return B2Q_ERR_ARG;
    AbsmaxSrc am{w->absmax, w->absmax_q, w->absmax2, w->code256, w->offset};
    static int cfg = -1;   // tuning hook: B2Q_GEMV_CFG = 0 (2 rows/warp, 2 in flight), 1 (1 row, 4 in flight, 8 blocks/SM), 2 (2 rows, 2, 6 blocks/SM), 3 (tensor-core formulation, default)
    if (cfg < 0) { const char* e = getenv("B2Q_GEMV_CFG"); cfg = e ? atoi(e) : 3; }
    const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_bf16);
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(y_bf16);
    const uint4* pk = reinterpret_cast<const uint4*>(w->packed);
#define B2Q_GEMV_LAUNCH(MR, RPW_, U_, MINB_) \
    nf4_gemv_kernel<MR, RPW_, U_, MINB_><<<(N + 4 * RPW_ - 1) / (4 * RPW_), 128, 0, stream>>>(x, pk, am, w->code16, y, M, N, K)
    if (cfg == 3) {
        nf4_gemv_mma_kernel<4><<<(N + 15) / 16, 256, 0, stream>>>(x, pk, am, w->code16, y, M, N, K);
    } else if (M == 1) {
        if (cfg == 0) B2Q_GEMV_LAUNCH(1, 2, 2, 1); else if (cfg == 2) B2Q_GEMV_LAUNCH(1, 2, 2, 6); else B2Q_GEMV_LAUNCH(1, 1, 4, 8);
    } else if (M <= 4) {
        if (cfg == 0) B2Q_GEMV_LAUNCH(4, 2, 2, 1); else B2Q_GEMV_LAUNCH(4, 2, 2, 4);
    } else {
        if (cfg == 0) B2Q_GEMV_LAUNCH(8, 2, 2, 1); else B2Q_GEMV_LAUNCH(8, 2, 2, 4);
    }
#undef B2Q_GEMV_LAUNCH
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int b2q_nf4_quantize(const void* w, int w_is_bf16, int64_t n, uint8_t* packed, float* absmax,
                                cudaStream_t stream) {
    if (n == 0) return 0;
    if (n % 2 != 0) return B2Q_ERR_SHAPE;
    const long long nblocks = (n + 63) / 64;
    const int grid = grid_for(nblocks * 32, 256);
    if (w_is_bf16)
        nf4_quantize_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(w),
                                                                     packed, absmax, n);
    else
        nf4_quantize_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(w), packed, absmax, n);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int b2q_absmax_double_quant(const float* absmax, int64_t nblocks, const float* code256, uint8_t* absmax_q,
                                       float* absmax2, float* offset_out, cudaStream_t stream) {
    if (nblocks == 0) return 0;
    absmax_mean_kernel<<<1, 1024, 0, stream>>>(absmax, nblocks, offset_out);
    const int grid = static_cast<int>((nblocks + 255) / 256);
    absmax_quantize_kernel<<<grid, 256, 0, stream>>>(absmax, offset_out, code256, absmax_q, absmax2, nblocks);
    count_launch(2);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int b2q_dropout_mask(uint8_t* mask, int64_t n, uint64_t seed, float p, cudaStream_t stream) {
    if (n == 0) return 0;
    if (!(p >= 0.f && p < 1.f)) return B2Q_ERR_ARG;
    dropout_mask_kernel<<<grid_for(n, 256), 256, 0, stream>>>(mask, n, seed, dropout_threshold(p));
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int b2q_dropout_apply(const void* x_bf16, void* xd_bf16, int64_t n, uint64_t seed, float p,
                                 cudaStream_t stream) {
    if (n == 0) return 0;
    if (n % 8 != 0) return B2Q_ERR_SHAPE;
    if (!(p >= 0.f && p < 1.f)) return B2Q_ERR_ARG;
    dropout_apply_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x_bf16),
                                                                   reinterpret_cast<uint4*>(xd_bf16), n / 8, seed,
                                                                   dropout_threshold(p), 1.0f / (1.0f - p));
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int b2q_reduce_partials(const float* partial, int splits, int64_t n, float scale, void* out_bf16,
                                   int accumulate, cudaStream_t stream) {
    if (n == 0) return 0;
    reduce_partials_kernel<<<grid_for(n, 256), 256, 0, stream>>>(partial, splits, n, scale,
                                                                 reinterpret_cast<__nv_bfloat16*>(out_bf16),
                                                                 accumulate);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int b2q_dropout_bwd_add(void* dx_bf16, const void* dxl_bf16, int64_t n, uint64_t seed, float p,
                                   cudaStream_t stream) {
    if (n == 0) return 0;
    if (n % 8 != 0) return B2Q_ERR_SHAPE;
    if (!(p >= 0.f && p < 1.f)) return B2Q_ERR_ARG;
    dropout_bwd_add_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<uint4*>(dx_bf16),
                                                                     reinterpret_cast<const uint4*>(dxl_bf16), n / 8,
                                                                     seed, dropout_threshold(p), 1.0f / (1.0f - p));
    count_launch();
    return static_cast<int>(cudaGetLastError());
}
