// Internal helpers shared by the .cu files of libb2q.so.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/b2q.h"

namespace b2q {

// Counter-based keep/drop decision: 64-bit mix of (seed, index) -> 32 bits, keep iff >= thresh.
__host__ __device__ __forceinline__ uint32_t dropout_hash(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = idx + seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return static_cast<uint32_t>(z >> 32);
}
__host__ __device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long idx,
                                                      uint32_t thresh) {
    return dropout_hash(seed, idx) >= thresh;
}
inline uint32_t dropout_threshold(float p) {
    const double t = static_cast<double>(p) * 4294967296.0;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(t);
}

extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

}  // namespace b2q
