// Fused bf16 AdamW + global-norm gradient clipping over the flat LoRA buckets.
// Replaces, for the trainable LoRA parameters, what the reference does with
//   accel.clip_grad_norm_(model.parameters(), GRAD_MAX)        pipeline/CuLLaVOPipeline.py:90-91
//   torch.optim.AdamW(...).step()                              trainer/cullavo_trainer.py:13, trainer/default_trainer.py:86-90
// i.e. one multi-tensor norm + ~10 foreach kernels over 448 tensors.  HBM-bound: 14 bytes per element
// (read p, g, m, v; write p, m, v), 128-bit accesses, grid = a multiple of the SM count.
#include <cuda_bf16.h>

#include "b2q_internal.h"

namespace b2q {

constexpr int OPT_THREADS = 256;
constexpr int OPT_MAX_BLOCKS = 148 * 8;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
    if (w == 0) {
        t = l < OPT_THREADS / 32 ? red[l] : 0.f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) red[0] = t;
    }
    __syncthreads();
    t = red[0];
    __syncthreads();
    return t;
}

// partials[b] = sum over this block's elements of g^2 (fp32); fixed assignment of elements to blocks and a fixed
// reduction tree, so the norm is bit-reproducible run to run.
__global__ void __launch_bounds__(OPT_THREADS) sqnorm_partials_kernel(const uint4* __restrict__ g, long long n8,
                                                                     const __nv_bfloat16* __restrict__ tail, int ntail,
                                                                     float* __restrict__ partials) {
    __shared__ float red[OPT_THREADS / 32];
    float acc = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(OPT_THREADS) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * OPT_THREADS) {
        const uint4 q = __ldg(g + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
            const float lo = __bfloat162float(b.x), hi = __bfloat162float(b.y);
            acc = fmaf(lo, lo, acc);
            acc = fmaf(hi, hi, acc);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < ntail) {
        const float t = __bfloat162float(tail[threadIdx.x]);
        acc = fmaf(t, t, acc);
    }
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

struct AdamArgs {
    float lr, beta1, beta2, eps, weight_decay, inv_bc1, inv_sqrt_bc2, max_norm;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamArgs& a, float clip) {
    g *= clip;
    p *= (1.0f - a.lr * a.weight_decay);                  // decoupled weight decay (AdamW)
    m = a.beta1 * m + (1.0f - a.beta1) * g;
    v = a.beta2 * v + (1.0f - a.beta2) * g * g;
    const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
    p -= (a.lr * a.inv_bc1) * (m / denom);
}

template <bool STATE_F32>
__global__ void __launch_bounds__(OPT_THREADS) adamw_kernel(__nv_bfloat16* __restrict__ p, const __nv_bfloat16* __restrict__ g,
                                                            void* __restrict__ m_, void* __restrict__ v_, long long n,
                                                            AdamArgs a, const float* __restrict__ partials, int n_partials) {
    __shared__ float red[OPT_THREADS / 32];
    float clip = 1.0f;
    if (a.max_norm > 0.f && partials != nullptr) {
        float acc = 0.f;
        for (int i = threadIdx.x; i < n_partials; i += OPT_THREADS) acc += __ldg(partials + i);
        const float total = block_sum(acc, red);            // same order in every block -> same clip everywhere
        const float c = a.max_norm / (sqrtf(total) + 1e-6f);  // torch.nn.utils.clip_grad_norm_
        clip = c < 1.0f ? c : 1.0f;
    }
    const long long n8 = n / 8;
    for (long long i = blockIdx.x * static_cast<long long>(OPT_THREADS) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * OPT_THREADS) {
        uint4 pq = reinterpret_cast<uint4*>(p)[i];
        const uint4 gq = __ldg(reinterpret_cast<const uint4*>(g) + i);
        uint32_t pw[4] = {pq.x, pq.y, pq.z, pq.w};
        const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w};
        if constexpr (STATE_F32) {
            float4* m4 = reinterpret_cast<float4*>(m_) + 2 * i;
            float4* v4 = reinterpret_cast<float4*>(v_) + 2 * i;
            float4 ma = m4[0], mb = m4[1], va = v4[0], vb = v4[1];
            float mm[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
            float vv[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 pb = *reinterpret_cast<__nv_bfloat162*>(&pw[j]);
                const __nv_bfloat162 gb = *reinterpret_cast<const __nv_bfloat162*>(&gw[j]);
                float p0 = __bfloat162float(pb.x), p1 = __bfloat162float(pb.y);
                adam_elem(p0, __bfloat162float(gb.x), mm[2 * j], vv[2 * j], a, clip);
                adam_elem(p1, __bfloat162float(gb.y), mm[2 * j + 1], vv[2 * j + 1], a, clip);
                pb = __floats2bfloat162_rn(p0, p1);
                pw[j] = *reinterpret_cast<uint32_t*>(&pb);
            }
            m4[0] = make_float4(mm[0], mm[1], mm[2], mm[3]); m4[1] = make_float4(mm[4], mm[5], mm[6], mm[7]);
            v4[0] = make_float4(vv[0], vv[1], vv[2], vv[3]); v4[1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
        } else {
            uint4 mq = reinterpret_cast<uint4*>(m_)[i], vq = reinterpret_cast<uint4*>(v_)[i];
            uint32_t mw[4] = {mq.x, mq.y, mq.z, mq.w}, vw[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 pb = *reinterpret_cast<__nv_bfloat162*>(&pw[j]);
                const __nv_bfloat162 gb = *reinterpret_cast<const __nv_bfloat162*>(&gw[j]);
                __nv_bfloat162 mb = *reinterpret_cast<__nv_bfloat162*>(&mw[j]);
                __nv_bfloat162 vb = *reinterpret_cast<__nv_bfloat162*>(&vw[j]);
                float p0 = __bfloat162float(pb.x), p1 = __bfloat162float(pb.y);
                float m0 = __bfloat162float(mb.x), m1 = __bfloat162float(mb.y);
                float v0 = __bfloat162float(vb.x), v1 = __bfloat162float(vb.y);
                adam_elem(p0, __bfloat162float(gb.x), m0, v0, a, clip);
                adam_elem(p1, __bfloat162float(gb.y), m1, v1, a, clip);
                pb = __floats2bfloat162_rn(p0, p1);
                mb = __floats2bfloat162_rn(m0, m1);
                vb = __floats2bfloat162_rn(v0, v1);
                pw[j] = *reinterpret_cast<uint32_t*>(&pb);
                mw[j] = *reinterpret_cast<uint32_t*>(&mb);
                vw[j] = *reinterpret_cast<uint32_t*>(&vb);
            }
            reinterpret_cast<uint4*>(m_)[i] = make_uint4(mw[0], mw[1], mw[2], mw[3]);
            reinterpret_cast<uint4*>(v_)[i] = make_uint4(vw[0], vw[1], vw[2], vw[3]);
        }
        reinterpret_cast<uint4*>(p)[i] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
    }
    // tail (n % 8 elements), one thread each
    const long long t0 = n8 * 8;
    if (blockIdx.x == 0 && t0 + threadIdx.x < n) {
        const long long i = t0 + threadIdx.x;
        float pv = __bfloat162float(p[i]);
        float mv, vv;
        if constexpr (STATE_F32) { mv = static_cast<float*>(m_)[i]; vv = static_cast<float*>(v_)[i]; }
        else { mv = __bfloat162float(static_cast<__nv_bfloat16*>(m_)[i]); vv = __bfloat162float(static_cast<__nv_bfloat16*>(v_)[i]); }
        adam_elem(pv, __bfloat162float(g[i]), mv, vv, a, clip);
        p[i] = __float2bfloat16_rn(pv);
        if constexpr (STATE_F32) { static_cast<float*>(m_)[i] = mv; static_cast<float*>(v_)[i] = vv; }
        else { static_cast<__nv_bfloat16*>(m_)[i] = __float2bfloat16_rn(mv); static_cast<__nv_bfloat16*>(v_)[i] = __float2bfloat16_rn(vv); }
    }
}

static int opt_blocks(long long n) {
    long long b = (n / 8 + OPT_THREADS - 1) / OPT_THREADS;
    if (b < 1) b = 1;
    if (b > OPT_MAX_BLOCKS) b = OPT_MAX_BLOCKS;
    return static_cast<int>(b);
}

}  // namespace b2q

using namespace b2q;

extern "C" int b2q_sqnorm_blocks(int64_t n) { return n <= 0 ? 0 : opt_blocks(n); }

extern "C" int b2q_sqnorm_partials(const void* g_bf16, int64_t n, float* partials, cudaStream_t stream) {
    if (n <= 0) return 0;
    if (g_bf16 == nullptr || partials == nullptr) return B2Q_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(g_bf16) & 15) != 0) return B2Q_ERR_ARG;
    const long long n8 = n / 8;
    sqnorm_partials_kernel<<<opt_blocks(n), OPT_THREADS, 0, stream>>>(
        static_cast<const uint4*>(g_bf16), n8, static_cast<const __nv_bfloat16*>(g_bf16) + n8 * 8, static_cast<int>(n - n8 * 8),
        partials);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int b2q_adamw_step(void* p_bf16, const void* g_bf16, void* m, void* v, int state_is_f32, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int64_t step,
                              const float* sq_partials, int n_partials, float max_norm, cudaStream_t stream) {
    if (n <= 0) return 0;
    if (p_bf16 == nullptr || g_bf16 == nullptr || m == nullptr || v == nullptr || step < 1) return B2Q_ERR_ARG;
    if (((reinterpret_cast<uintptr_t>(p_bf16) | reinterpret_cast<uintptr_t>(g_bf16) | reinterpret_cast<uintptr_t>(m) |
          reinterpret_cast<uintptr_t>(v)) & 15) != 0)
        return B2Q_ERR_ARG;
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
    a.inv_bc1 = static_cast<float>(1.0 / bc1);
    a.inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
    if (state_is_f32)
        adamw_kernel<true><<<opt_blocks(n), OPT_THREADS, 0, stream>>>(static_cast<__nv_bfloat16*>(p_bf16),
            static_cast<const __nv_bfloat16*>(g_bf16), m, v, n, a, sq_partials, n_partials);
    else
        adamw_kernel<false><<<opt_blocks(n), OPT_THREADS, 0, stream>>>(static_cast<__nv_bfloat16*>(p_bf16),
            static_cast<const __nv_bfloat16*>(g_bf16), m, v, n, a, sq_partials, n_partials);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}
