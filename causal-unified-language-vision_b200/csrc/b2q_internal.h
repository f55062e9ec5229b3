// Internal helpers shared by the .cu files of libb2q.so.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/b2q.h"

namespace b2q {

// Counter-based keep/drop decision shared by every kernel that needs the LoRA-dropout mask, by the
// stand-alone mask kernel and (through it) by the CPU oracle.  One 32-bit hash of (seed, i >> 1) serves
// two neighbouring elements, 16 bits each: keep(i) = r16(i) >= thresh16, thresh16 = round(p * 65536).
// The effective drop rate is thresh16 / 65536 (p = 0.05 -> 0.0500031); the survivors are scaled by
// 1 / (1 - p) exactly as torch's dropout does.
__host__ __device__ __forceinline__ uint32_t dropout_hash32(uint32_t seed_lo, uint32_t seed_hi, uint32_t j) {
    uint32_t x = (j * 0x9E3779B1u) ^ seed_lo;
    x ^= x >> 16;
    x *= 0x85EBCA6Bu;
    x ^= (x >> 13) ^ seed_hi;
    x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx,
                                                      uint32_t thresh16) {
    const uint32_t h = dropout_hash32(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32),
                                      static_cast<uint32_t>(idx >> 1));
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
