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
// Packed form of the same mask (experiment: consumers read bits instead of hashing; B2Q_MASK_BITS=1).  Bit k of byte b is
// keep(8 b + k).  dropout_bits32 produces the 32 bits of elements [e0, e0 + 32) (e0 a multiple of 32) from 8 hashes.
__host__ __device__ __forceinline__ uint32_t dropout_bits32(uint32_t seed_lo, uint32_t seed_hi, unsigned long long e0,
                                                            uint32_t thr2) {
    const uint32_t thr = thr2 & 0xFFFFu;
    const uint32_t j0 = static_cast<uint32_t>(e0 >> 2);
    uint32_t bits = 0;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        uint32_t a, b;
        dropout_hash64(seed_lo, seed_hi, j0 + h, a, b);
        const uint32_t k = ((a & 0x7FFFu) >= thr ? 1u : 0u) | (((a >> 16) & 0x7FFFu) >= thr ? 2u : 0u) |
                           ((b & 0x7FFFu) >= thr ? 4u : 0u) | (((b >> 16) & 0x7FFFu) >= thr ? 8u : 0u);
        bits |= k << (4 * h);
    }
    return bits;
}
// One mask byte (8 consecutive elements = one 16-byte chunk of bf16) -> the four bf16x2 AND-masks of the chunk's words.
// Two multiplies move bit i of each nibble to the sign bit of byte i (the partial products never share a bit position,
// so nothing carries); PRMT in sign-replicate mode then expands byte signs to byte masks, two bytes per element.
__host__ __device__ __forceinline__ void dropout_byte_to_masks(uint32_t byte, uint32_t (&m)[4]) {
    const uint32_t x0 = (byte & 0xFu) * 0x10204080u;          // bits 0..3 -> signs of bytes 0..3
    const uint32_t x1 = ((byte >> 4) & 0xFu) * 0x10204080u;   // bits 4..7
#ifdef __CUDA_ARCH__
    asm("prmt.b32 %0, %1, %1, 0x9988;" : "=r"(m[0]) : "r"(x0));
    asm("prmt.b32 %0, %1, %1, 0xBBAA;" : "=r"(m[1]) : "r"(x0));
    asm("prmt.b32 %0, %1, %1, 0x9988;" : "=r"(m[2]) : "r"(x1));
    asm("prmt.b32 %0, %1, %1, 0xBBAA;" : "=r"(m[3]) : "r"(x1));
#else
    const uint32_t xs[2] = {x0, x1};
    for (int j = 0; j < 4; ++j) {
        const uint32_t x = xs[j >> 1];
        const int lo = (j & 1) ? 23 : 7, hi = (j & 1) ? 31 : 15;   // sign bits of bytes (0,1) or (2,3)
        m[j] = (((x >> lo) & 1u) ? 0x0000FFFFu : 0u) | (((x >> hi) & 1u) ? 0xFFFF0000u : 0u);
    }
#endif
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
