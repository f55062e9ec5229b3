// Internal helpers shared by the .cu files of libb2q.so.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/b2q.h"

namespace b2q {

// Counter-based keep/drop decision shared by every kernel that needs the LoRA-dropout mask, by the stand-alone
// mask kernel and (through it) by the CPU oracle.  One call of a 4-round Philox-2x32 style mixer on
// (counter = i >> 2, key = seed) yields 64 bits = four 16-bit fields for four neighbouring elements:
// keep(i) = field(i & 3) >= thresh16, thresh16 = round(p * 65536).  Two IMAD.WIDE + LOP3 per round, i.e. ~2
// integer operations per element (the previous 3-multiply hash cost ~5 and made the in-shared-memory dropout of
// the LoRA GEMMs issue-bound).  The effective drop rate is thresh16 / 65536 (p = 0.05 -> 0.0500031); the survivors
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
// 32-bit AND-mask for the two bf16 elements that share one hash word (low field = even element)
__host__ __device__ __forceinline__ uint32_t dropout_mask2(uint32_t h, uint32_t thresh16) {
    return ((h & 0xFFFFu) >= thresh16 ? 0x0000FFFFu : 0u) | ((h >> 16) >= thresh16 ? 0xFFFF0000u : 0u);
}
__host__ __device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx,
                                                      uint32_t thresh16) {
    uint32_t a, b;
    dropout_hash64(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), static_cast<uint32_t>(idx >> 2), a, b);
    const uint32_t h = (idx & 2ull) ? b : a;
    const uint32_t r = (idx & 1ull) ? (h >> 16) : (h & 0xFFFFu);
    return r >= thresh16;
}
inline uint32_t dropout_threshold(float p) {
    const double t = static_cast<double>(p) * 65536.0 + 0.5;
    return t >= 65535.0 ? 65535u : static_cast<uint32_t>(t);
}

extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

}  // namespace b2q
