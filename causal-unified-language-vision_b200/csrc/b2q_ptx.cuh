// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA, tcgen05 (UMMA / TMEM), clusters.
// Hand-written for this repo; no CUTLASS / CuTe types.  Everything here is device-only.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace b2q {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xFFFFFFFF;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(pred));
    return pred;
}

// ---------------------------------------------------------------- cluster ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync() {
    cluster_arrive();
    cluster_wait();
}
// shared::cta address of this CTA  ->  shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

// --------------------------------------------------------------- mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Scopes.  Every wait / arrive uses the PTX defaults (.acquire / .release at .cta scope), also the arrives that the peer
// CTA of a cta_group::2 pair sends to the LEADER's full / "accumulator drained" barriers (`_xcta` forms below).  Under
// the letter of the PTX memory model a .release.cta in CTA 1 does not synchronise-with an .acquire.cta in CTA 0; what
// the protocol needs is narrower and holds physically: the data the peer publishes lives in ITS OWN shared memory
// and is consumed by the tensor core through the async proxy, the publishing warp executes MEMBAR.ALL.CTA +
// FENCE.VIEW.ASYNC (fence.proxy.async) before its arrive is issued, and the leader's issuing thread observes the phase
// flip before it issues the MMA -- no cache is involved on either side.  CUTLASS's ClusterBarrier::arrive(cta_id) is the
// same unscoped form.  Round 2 measured the formally scoped variant (B2Q_XCTA_SCOPE_CLUSTER=1: .release.cluster arrives,
// .acquire.cluster waits; ptxas adds MEMBAR.ALL.GPU / CCTL.IVALL per stage): whole step 35.9k -> 28.4k tokens/s
// (-21 %), and it changed nothing about the two real bugs of the round, so .cta stays the default
// and the scoped build stays available for A/B runs.
#ifndef B2Q_XCTA_SCOPE_CLUSTER
#define B2Q_XCTA_SCOPE_CLUSTER 0
#endif
#define B2Q_XCTA_SCOPE_CTA (!B2Q_XCTA_SCOPE_CLUSTER)
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// `bar` is a shared::cluster address (own CTA's shared::cta addresses are valid ones); may name the peer's barrier
__device__ __forceinline__ void mbar_arrive_xcta(uint32_t bar) {
#if B2Q_XCTA_SCOPE_CTA
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
#else
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_xcta(uint32_t bar, uint32_t bytes) {
#if B2Q_XCTA_SCOPE_CTA
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#else
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
#endif
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// wait on a barrier of this CTA that threads of the peer CTA arrive on
__device__ __forceinline__ uint32_t mbar_try_wait_xcta(uint32_t bar, uint32_t parity) {
#if B2Q_XCTA_SCOPE_CTA
    return mbar_try_wait(bar, parity);
#else
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
#endif
}

// ------------------------------------------------------------ stall guard ----
// Every pipeline wait is bounded.  A wait that has not been satisfied after STALL_LIMIT_CYCLES (about 2-3 s at
// 1.3-2.0 GHz; the longest legitimate wait in these kernels is one output tile, < 0.1 ms) writes a record -- kernel
// configuration, CTA, thread, wait site, barrier index and parity, tile / k-block, launch geometry and the raw 64-bit
// state of every barrier of the CTA -- to a host-mapped buffer (StallSink::buf, owned by the host side of the library,
// readable after the context is dead) and, a quarter of the limit later (so that the other stuck roles of the CTA get
// their records out too), traps.  A lost arrive or a wrong parity therefore ends the process with a report
// (b2q_debug_stall_report) instead of spinning until an external watchdog kills the job.
#ifndef B2Q_STALL_GUARD
#define B2Q_STALL_GUARD 1
#endif
constexpr long long STALL_LIMIT_CYCLES = 4000000000ll;
constexpr uint32_t STALL_MAGIC = 0xB2517A11u;
constexpr int STALL_MAX_RECORDS = 96;
constexpr int STALL_MAX_BARS = 48;
constexpr int STALL_HDR_WORDS = 16;                              // [0] record counter, [1] magic, rest reserved
constexpr int STALL_REC_WORDS = 24 + 2 * STALL_MAX_BARS;         // header of the record + barrier words
constexpr size_t STALL_BUF_BYTES = 4ull * (STALL_HDR_WORDS + STALL_MAX_RECORDS * STALL_REC_WORDS);

struct StallSink {
    uint32_t* buf;       // host-mapped record buffer (nullptr: trap without a record)
    uint32_t bar_base;   // shared::cta address of barrier 0 of this CTA
    uint32_t nbars;
    uint32_t cfg;        // GemmCfg::ID
    uint32_t geom[6];    // launch geometry: M, N, kb_main, kb_tail, splits, tiles
    long long t0[32];    // per warp: clock64 at the first check of the wait in progress (scratch of stall_check)
};

__device__ __noinline__ void stall_record(const StallSink* sk, uint32_t bar, uint32_t parity, uint32_t site, int a,
                                          int b) {
    uint32_t* buf = sk->buf;
    if (buf == nullptr) return;
    const unsigned act = __activemask();
    if ((threadIdx.x & 31u) != static_cast<unsigned>(__ffs(act) - 1)) return;   // one record per warp
    const uint32_t slot = atomicAdd_system(buf, 1u);
    if (slot >= static_cast<uint32_t>(STALL_MAX_RECORDS)) return;
    volatile uint32_t* r = buf + STALL_HDR_WORDS + slot * STALL_REC_WORDS;
    uint32_t smid, crank;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    r[1] = sk->cfg; r[2] = blockIdx.x; r[3] = threadIdx.x; r[4] = site; r[5] = (bar - sk->bar_base) >> 3;
    r[6] = parity; r[7] = static_cast<uint32_t>(a); r[8] = static_cast<uint32_t>(b); r[9] = crank; r[10] = sk->nbars;
    r[11] = smid; r[12] = act; r[13] = gridDim.x;
#pragma unroll
    for (int i = 0; i < 6; ++i) r[14 + i] = sk->geom[i];
    const uint32_t nb = sk->nbars < static_cast<uint32_t>(STALL_MAX_BARS) ? sk->nbars : static_cast<uint32_t>(STALL_MAX_BARS);
    for (uint32_t i = 0; i < nb; ++i) {
        uint32_t lo, hi;
        asm volatile("ld.volatile.shared::cta.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(sk->bar_base + 8u * i));
        r[24 + 2 * i] = lo;
        r[25 + 2 * i] = hi;
    }
    __threadfence_system();
    r[0] = STALL_MAGIC;   // record complete
    __threadfence_system();
}

// Cold part of a bounded wait: called every 4096 failed probes with the probe count `n` (bit 31: record written).
// The start time lives in the sink (one slot per warp: the lanes of a warp wait together), so the hot loop keeps a single
// counter register.  Returns the new n.
__device__ __noinline__ uint32_t stall_check(StallSink* sk, uint32_t n, uint32_t bar, uint32_t parity, uint32_t site,
                                             int a, int b) {
    const long long now = clock64();
    volatile long long* t0 = &sk->t0[threadIdx.x >> 5];
    if ((n & 0x7FFFFFFFu) == 0x1000u) { *t0 = now; return n; }
    const long long dt = now - *t0;
    if (dt > STALL_LIMIT_CYCLES && (n >> 31) == 0u) {
        stall_record(sk, bar, parity, site, a, b);
        n |= 0x80000000u;
    }
    if (dt > STALL_LIMIT_CYCLES + STALL_LIMIT_CYCLES / 4) __trap();
    return n;
}

template <bool XCTA = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, const StallSink* sk, uint32_t site, int a = 0,
                                          int b = 0) {
#if B2Q_STALL_GUARD
    uint32_t n = 0;
    while (!(XCTA ? mbar_try_wait_xcta(bar, parity) : mbar_try_wait(bar, parity))) {
        if (((++n) & 0xFFFu) == 0u) n = stall_check(const_cast<StallSink*>(sk), n, bar, parity, site, a, b);
    }
#else
    while (!(XCTA ? mbar_try_wait_xcta(bar, parity) : mbar_try_wait(bar, parity))) {
    }
#endif
}

// -------------------------------------------------------------------- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// 2D tiled load, completion on a barrier of the executing CTA.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// 2D tiled load issued inside a CTA pair; `bar` is a shared::cluster address and may
// name the peer CTA's barrier (this is what .cta_group::2 permits).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// L2 prefetch of a 2D tile (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_2d(const void* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
// D[tile] += smem tile, element-wise in the tensor map's data type (bf16), performed at the L2
__device__ __forceinline__ void tma_reduce_add_2d(const void* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- tcgen05 ----
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                     : "memory");
    else
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                     : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
    if constexpr (CG == 1)
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    else
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc];  kind::f16 covers bf16 inputs with fp32 accumulate.
template <int CG>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
    if constexpr (CG == 1) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(d_tmem),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// Arrive on `bar` once every previously issued tcgen05.mma of this thread has completed.
// CG==2: the arrive is multicast to the same barrier offset in both CTAs of the pair.
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                     : "memory");
    } else {
        const uint16_t mask = 0x3;
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                bar),
            "h"(mask)
            : "memory");
    }
}

// TMEM -> registers: 32 lanes x 32 columns of 32 bit; thread i of the warp gets lane i.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------ descriptors ----
// Shared-memory matrix descriptor (64 bit) for tcgen05.mma, SWIZZLE_128B layouts.
//   bits [0,14)  start address >> 4          bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset >> 4 bits [46,48) version = 1 (Blackwell)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
// K-major  : 8-row groups of 128 B rows, SBO = byte pitch between 8-row groups (LBO unused = 1).
// MN-major : 64-element (128 B) chunks along MN, LBO = pitch between chunks along MN,
//            SBO = pitch between 8-deep groups along K.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// Instruction descriptor (32 bit) for kind::f16 with bf16 A/B and fp32 D.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt (1 = bf16)
//   [15] A major (1 = MN)  [16] B major (1 = MN)    [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) |
           (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

}  // namespace b2q
