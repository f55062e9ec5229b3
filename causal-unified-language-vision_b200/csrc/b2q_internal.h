// Internal helpers shared by the .cu files of libb2q.so.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/b2q.h"

namespace b2q {

// Text behind b2q_last_error_detail() (defined in qlora_gemm.cu).
void set_error_detail(const char* msg);

// Counter-based keep/drop decision shared by every kernel that needs the LoRA-dropout mask, by the stand-alone
// mask kernel and (through it) by the CPU oracle.  One call of a 4-round Philox-2x32 style mixer on
// (counter = i >> 2, key = seed) yields 64 bits = four 15-bit fields (the low 15 bits of each 16-bit quarter) for four
// neighbouring elements: keep(i) = field(i & 3) >= thr, thr = round(p * 32768).  Two IMAD.WIDE + LOP3 per round, i.e. ~2
// integer operations per element (the previous 3-multiply hash cost ~5 and made the in-shared-memory dropout of
// the LoRA GEMMs issue-bound).  The effective drop rate is thr / 32768 (p = 0.05 -> 0.0499878); the survivors
// are scaled by 1 / (1 - p) exactly as torch's dropout does.  Torch's own Philox stream is NOT reproduced.
__host__ __device__ __forceinline__ void dropout_hash64(uint32_t seed_lo, uint32_t seed_hi, uint32_t j, uint32_t& a,
                                                        uint32_t& b) {
    a = j;
    b = seed_hi;
    uint32_t key = seed_lo;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned long long pr = static_cast<unsigned long long>(a) * 0xD2511F53ull;
        a = static_cast<uint32_t>(pr >> 32) ^ b ^ key;
        b = static_cast<uint32_t>(pr);
        key += 0x9E3779B9u;
    }
}
// 32-bit AND-mask for the two bf16 elements that share one hash word (low half = even element).  A field is the low
// 15 bits of a 16-bit half; `thr2` holds the 15-bit threshold in both halves (dropout_threshold()).  With bit 15 of
// each half forced to 1 the two subtractions cannot borrow into each other and bit 15 of each half of the difference
// says field >= threshold; PRMT in sign-replicate mode expands those two bits to byte masks: 3 integer operations
// instead of 2 compares + 2 selects + 1 or (the in-shared-memory dropout pass is bound by exactly these).
__host__ __device__ __forceinline__ uint32_t dropout_mask2(uint32_t h, uint32_t thr2) {
    const uint32_t t = ((h & 0x7FFF7FFFu) | 0x80008000u) - thr2;
#ifdef __CUDA_ARCH__
    uint32_t m;
    asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(t));
    return m;
#else
    return ((t & 0x8000u) ? 0x0000FFFFu : 0u) | ((t & 0x80000000u) ? 0xFFFF0000u : 0u);
#endif
}
__host__ __device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx, uint32_t thr2) {
    uint32_t a, b;
    dropout_hash64(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), static_cast<uint32_t>(idx >> 2), a, b);
    const uint32_t h = (idx & 2ull) ? b : a;
    const uint32_t r = ((idx & 1ull) ? (h >> 16) : h) & 0x7FFFu;
    return r >= (thr2 & 0xFFFFu);
}
// 15-bit threshold round(p * 32768), replicated into both 16-bit halves
inline uint32_t dropout_threshold(float p) {
    const double t = static_cast<double>(p) * 32768.0 + 0.5;
    const uint32_t v = t >= 32767.0 ? 32767u : static_cast<uint32_t>(t);
    return v * 0x00010001u;
}

extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

}  // namespace b2q
