// Warp-specialised, persistent tcgen05 GEMM for sm_100a with an optional in-main-loop NF4
// decode of the B operand.  One template, instantiated for every GEMM on the QLoRA path:
//
//   D[M,N] = sum over main k-blocks  A[M, 64] * Bop[64, N]      (Bop: bf16 via TMA, or NF4 decoded)
//          + sum over tail k-blocks  A2[M, 64] * B2[64, N]      (LoRA tail, bf16 via TMA)
//
// Roles (one CTA per SM, or a CTA pair with cta_group::2); warp ids are GemmCfg::W_*:
//   TMA producer (A, TMA-fed B)      UMMA issuer (one thread of the leader CTA)      TMEM allocator
//   TMA producer of the packed-NF4 ring (B_DEC)
//   warps 4-7   epilogue (TMEM -> registers -> shared-memory staging -> global)
//   NF4 decode (B_DEC): 8 warps = NG groups of BNC threads taking k-blocks round-robin:
//               packed bytes (TMA-staged smem) -> registers (LUT/PRMT) -> bf16 in the canonical
//               128B-swizzled UMMA operand layout in smem
//   warps 8..   (A_XF) in-shared-memory LoRA-dropout transform of the A tile
// In the decode kernels the warp ids are chosen so that no decode warp shares a warp scheduler (id % 4) with the UMMA
// issuer: see the role map in GemmCfg.
//
// Shared memory: STAGES x {A tile (MT x 128 rows x 64 k bf16), B tile (BNC rows x 64 k bf16)}
// plus an independent, deeper ring of PST packed-NF4 tiles (BNC x 32 B) so the packed bytes are
// in shared memory long before the decoded-B slot they go to is free.  An operand stage is
// released by tcgen05.commit when the MMAs that read it have finished.  The bf16 weight only
// ever exists in shared memory.
//
// Operand layouts (both 128B-swizzled, see b2q_ptx.cuh umma_desc_sw128):
//   K-major  tile [rows x 64 k]: row r at (r/8)*1024 + (r%8)*128, 16-byte chunk c at (c ^ (r%8))*16
//   MN-major tile [64 k x mn]  : same image with "row" = contraction index and 64-element chunks
//                                of the MN dimension 8192 B apart (LBO), 8-row groups 1024 B (SBO)
// so one decode routine (one thread = one 64-weight quantisation block = one 128-byte row) serves
// the forward GEMM (W is K-major) and the dX GEMM (W is MN-major) alike.
#pragma once
#include "b2q_decode.cuh"
#include "b2q_internal.h"
#include "b2q_ptx.cuh"

namespace b2q {

enum EpiKind : int {
    EPI_BF16 = 0,      // D bf16 [M, ldd] = alpha * acc   (+ optional D2 = alpha2 * acc)
    EPI_F32_PART = 1,  // fp32 partial [split][M][N]
    EPI_F32_PART_T = 2,// fp32 partial, transposed [split][N][M]
    EPI_BF16_MASK = 3  // D bf16 [M, ldd] = keep(row*ldd+col) ? alpha * acc : 0   (LoRA-dropout backward)
};

struct GemmParams {
    CUtensorMap tmA;   // main A operand (bf16)
    CUtensorMap tmA2;  // tail A operand (bf16)
    CUtensorMap tmB;   // main B operand: bf16, or packed NF4 bytes when B_DEC
    CUtensorMap tmB2;  // tail B operand (bf16)
    CUtensorMap tmD;   // EPI_TMA: D as a bf16 tensor [M][ldd], box 64 cols x 32 rows, SWIZZLE_128B (store / reduce-add)
    AbsmaxSrc am;
    const float* code16;
    void* D;
    void* D2;
    long long ldd;     // row pitch of D / D2 in elements (EPI_BF16)
    float alpha, alpha2;
    int M, N;          // extents of D
    int kb_main;       // main k-blocks per tile (per split)
    int kb_tail;       // tail k-blocks per tile
    int splits;        // split-K factor (EPI_F32_PART*), else 1
    int kpr;           // absmax blocks per W row (= K_w / 64)          (B_DEC)
    int m_tiles, n_tiles;
    int group_m;       // rasterisation: m-tiles per L2 slab
    int reverse;       // walk the tiles in reverse order (a kernel that revisits the previous kernel's output starts with
                       // the part that is still in the L2)
    int tile_m;        // row stride between m-tiles, <= TILE_M (load-balancing of the skinny kernels: a tile still
                       // loads and multiplies TILE_M rows, but only stores its first tile_m -- the rest belong to the next tile)
    int accum_d;       // EPI_BF16: D = bf16(D + alpha * acc)  (read-modify-write of the caller's buffer)
    int stream_out;    // staged epilogue: store D with the streaming (evict-first) hint -- the output is not read again by the
                       // next kernel of the path, so it should not push the operands that ARE re-read out of the L2
    // LoRA dropout (A_XF transform / EPI_BF16_MASK): keep(i) for element i = row * xf_ld + col of the
    // activation the mask belongs to (b2q_internal.h dropout_keep)
    unsigned long long seed;
    unsigned int thresh16;
    unsigned int mask_flip;   // EPI_BF16_MASK: 0 = keep the kept elements, 0xFFFFFFFF = keep the DROPPED ones (dX correction)
    long long xf_ld;
    // optional phase trace (debug / profiling): clock64 stamps, [cta][tile][8]; nullptr = off
    long long* trace;
    int trace_tiles;
    int pf_dist;       // L2 prefetch distance of the A operand in k-blocks (0 = off)
    // stall guard (b2q_ptx.cuh): host-mapped record buffer of the library, or nullptr
    uint32_t* stall_buf;
};

template <int CG_, int MT_, int BN_, bool A_MN_, bool B_MN_, bool B_DEC_, int EPI_, int STAGES_, int NG_ = 2,
          int PST_ = 6, bool A_XF_ = false, int STG_ = 0, int ESETS_ = 1>
struct GemmCfg {
    // configuration id carried by stall records (decoded by b2q_debug_stall_report)
    static constexpr uint32_t ID = static_cast<uint32_t>(CG_) | (static_cast<uint32_t>(MT_) << 2) |
                                   (static_cast<uint32_t>(BN_ / 64) << 4) | (A_MN_ ? 1u << 8 : 0u) | (B_MN_ ? 1u << 9 : 0u) |
                                   (B_DEC_ ? 1u << 10 : 0u) | (static_cast<uint32_t>(EPI_) << 11) |
                                   (static_cast<uint32_t>(STAGES_) << 13) | (A_XF_ ? 1u << 17 : 0u) |
                                   (STG_ < 0 ? 1u << 18 : 0u) | (STG_ > 0 ? 1u << 19 : 0u) | (static_cast<uint32_t>(ESETS_ - 1) << 20);
    // ESETS = 2 or 3: that many sets of four epilogue warps (4-7, 8-11, 12-15), one per accumulator stage, for kernels that
    // are all epilogue (one k-block per tile): set e drains the tiles whose sequence number is e (mod ESETS).  (Round 2
    // measured 2 against 3 sets and the mask hash before / after the TMEM-load wait on the masked dX GEMM, same box: all
    // four within 0.3 % -- that kernel is bound by the HBM round trip of its read-modify-write of dX, not by its epilogue.)
    static constexpr int ESETS = ESETS_;
    // STG > 0: the bf16 epilogue goes TMEM -> registers -> swizzled shared-memory staging (32 rows x 64
    // columns per epilogue warp) -> TMA store (or TMA reduce-add when accum_d), so global memory sees whole
    // 128-byte row segments written by the copy engine instead of 32 scattered 16-byte stores per warp
    // instruction.
    // STG < 0: same staging idea without the copy engine: each epilogue warp transposes 32 rows x 64 columns
    // through a 4 KB swizzled tile and writes global memory itself, 4 whole 128-byte row segments per store
    // instruction (or read-modify-write when accum_d).
    static constexpr bool EPI_TMA = STG_ > 0;
    static constexpr bool EPI_COAL = STG_ < 0;
    static constexpr int STG_BYTES = STG_ != 0 ? ESETS_ * 4 * 4096 : 0;
    static constexpr bool A_XF = A_XF_;           // dropout mask applied to the A tile in shared memory
    static constexpr int XF_GROUPS = 4;                 // transform groups of 128 threads taking ring positions in turn
    static constexpr int XF_THREADS = A_XF_ ? 128 * XF_GROUPS : 0;
    static constexpr int NG = B_DEC_ ? NG_ : 0;   // decode groups (each BNC threads)
    static constexpr int PST = B_DEC_ ? PST_ : 0; // packed-NF4 ring depth
    static constexpr int CG = CG_;          // CTAs per MMA (cta_group)
    static constexpr int MT = MT_;          // 128-row accumulators per CTA
    static constexpr int BN = BN_;          // UMMA N (whole pair)
    static constexpr bool A_MN = A_MN_;
    static constexpr bool B_MN = B_MN_;
    static constexpr bool B_DEC = B_DEC_;
    static constexpr int EPI = EPI_;
    static constexpr int STAGES = STAGES_;

    static constexpr int BNC = BN / CG;                 // B rows held (and decoded) by one CTA
    static constexpr int TILE_M = 128 * MT * CG;        // rows of D per tile (pair)
    static constexpr int A_BYTES = MT * 128 * 64 * 2;
    static constexpr int B_BYTES = BNC * 64 * 2;
    static constexpr int P_BYTES = B_DEC ? BNC * 32 : 0;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES + PST * P_BYTES;
    static constexpr int ACC_COLS = MT * BN;
    static constexpr int ACC_STAGES = ESETS_ > 1 ? ESETS_ : ((2 * ACC_COLS <= 512) ? 2 : 1);
    static constexpr int TMEM_COLS_RAW = ACC_COLS * ACC_STAGES;
    static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                                     : TMEM_COLS_RAW <= 256 ? 256 : 512;
    static constexpr int NDT = B_DEC ? BNC : 0;         // decode threads per group
    static constexpr int NDW = NDT / 32;                // working decode warps per group (arrivals per stage)
    // Role -> warp id.  A warp's scheduler (SM sub-partition) is warp id % 4.  In the decode kernels every warp whose id
    // is 1 (mod 4) is a light or main-loop-idle role -- the single-thread UMMA issuer (1), an epilogue warp (5), the
    // operand TMA producer (9), the packed-ring producer + TMEM allocator (13) -- and the 8 decode warps take the ids
    // {0, 2, 3, 8, 10, 11, 12, 14}: ALU-heavy decode warps on the issuer's scheduler delay its instruction issue enough
    // to cost the main loop 13 % of its MMA rate (phase-trace experiment).
    static constexpr int W_PROD = B_DEC ? 9 : 0;
    static constexpr int W_MMA = 1;
    static constexpr int W_ALLOC = B_DEC ? 13 : 2;
    static constexpr int W_PACK = B_DEC ? 13 : -1;
    static constexpr int THREADS = 256 + NG * NDT + XF_THREADS + (ESETS_ - 1) * 128;
    static_assert(ESETS_ == 1 || ((ESETS_ == 2 || ESETS_ == 3) && !B_DEC_ && !A_XF_ && ESETS_ * MT_ * BN_ <= 512), "epilogue sets: one accumulator stage each, no decode / transform warps");
    static_assert(!B_DEC || NG * NDT == 256, "decode role map assumes 8 decode warps");
    static_assert(!(A_XF_ && B_DEC_), "A transform and B decode share the warps 8+");
    static_assert(!A_XF_ || (MT_ == 1 && CG_ == 1), "A transform: single 128-row tile, single CTA");
    static constexpr int BAR_BYTES = 1024;              // barriers + tmem ptr + code256
    static constexpr int CODE256_BYTES = B_DEC ? 1024 : 0;
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + RING_BYTES + BAR_BYTES + CODE256_BYTES + STG_BYTES;
    static_assert(STG_ == 0 || EPI_ == EPI_BF16 || EPI_ == EPI_BF16_MASK, "staged epilogues: bf16 output only");
    static_assert(STG_ == 0 || BN_ % 64 == 0, "staged epilogues work on 64-column groups");

    static_assert(BN % (16 * CG) == 0 && BN <= 256, "UMMA N");
    static_assert(ACC_COLS <= 512, "TMEM columns");
    static_assert(!B_DEC || (BNC % 64 == 0), "decode tile");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
    static_assert(NG == 0 || NG == 1 || NG == 2 || NG == 4, "decode groups");
    static_assert(THREADS <= 1024, "block size");
};

__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& mt, int& nt, int& split) {
    // split fastest, then m inside an L2 slab of group_m m-tiles, then n, then slab.
    if (p.reverse) tile = p.m_tiles * p.n_tiles * p.splits - 1 - tile;
    split = tile % p.splits;
    int t = tile / p.splits;
    const int per_slab = p.group_m * p.n_tiles;
    const int slab = t / per_slab;
    const int m_first = slab * p.group_m;
    const int m_cnt = min(p.group_m, p.m_tiles - m_first);
    const int within = t - slab * per_slab;
    nt = within / m_cnt;
    mt = m_first + within % m_cnt;
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) qlora_gemm_kernel(const __grid_constant__ GemmParams p) {
    constexpr int CG = Cfg::CG, MT = Cfg::MT, BN = Cfg::BN, BNC = Cfg::BNC, STAGES = Cfg::STAGES;
    constexpr int ACC_STAGES = Cfg::ACC_STAGES, ACC_COLS = Cfg::ACC_COLS;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024 B alignment
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    constexpr int PST = Cfg::PST, NG = Cfg::NG;
    const uint32_t bar_base = smem_base + Cfg::RING_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    // one "accumulator drained" barrier per (accumulator stage, 128-row sub-tile): the next tile's MMAs into
    // sub-tile 0 start while the epilogue is still draining sub-tile 1
    auto tempty_bar = [&](int a, int mt) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + a * MT + mt); };
    constexpr int NB0 = 2 * STAGES + ACC_STAGES + ACC_STAGES * MT;
    auto pk_bar = [&](int s) { return bar_base + 8u * (NB0 + s); };
    auto pk_empty_bar = [&](int s) { return bar_base + 8u * (NB0 + PST + s); };
    auto xf_bar = [&](int s) { return bar_base + 8u * (NB0 + 2 * PST + s); };
    constexpr int NBARS = NB0 + 2 * PST + (Cfg::A_XF ? STAGES : 0);
    // barrier area: NBARS barriers, TMEM base address (4 B), 4 B pad, 8 B scratch (sink of the decode warps' load fence),
    // stall-guard sink
    static_assert(8 * NBARS + 16 + sizeof(StallSink) <= Cfg::BAR_BYTES && NBARS <= STALL_MAX_BARS, "barrier area");
    const StallSink* sk = reinterpret_cast<const StallSink*>(smem_gen + Cfg::RING_BYTES + 8 * NBARS + 16);
    const uint32_t tmem_slot = bar_base + 8u * NBARS;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::RING_BYTES + 8 * NBARS);
    float* code256_s = reinterpret_cast<float*>(smem_gen + Cfg::RING_BYTES + Cfg::BAR_BYTES);

    auto a_stage = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
    auto b_stage = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
    auto p_stage = [&](int s) { return smem_base + STAGES * Cfg::STAGE_BYTES + s * Cfg::P_BYTES; };

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int pair_id = blockIdx.x / CG;
    const int num_pairs = gridDim.x / CG;
    const int num_tiles = p.m_tiles * p.n_tiles * p.splits;
    const int kb_total = p.kb_main + p.kb_tail;

    // ---------------------------------------------------------------- setup ----
    if (warp == Cfg::W_PROD && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        if (p.kb_tail > 0) {
            tma_prefetch_desc(&p.tmA2);
            tma_prefetch_desc(&p.tmB2);
        }
        if constexpr (Cfg::EPI_TMA) tma_prefetch_desc(&p.tmD);
    }
    if (warp == Cfg::W_MMA && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), CG * (1 + Cfg::NDW));  // per CTA: producer (+tx) + one arrive per decode warp
            mbar_init(empty_bar(s), 1);                   // tcgen05.commit
        }
        if constexpr (Cfg::A_XF)
            for (int s = 0; s < STAGES; ++s) mbar_init(xf_bar(s), 4);  // one arrive per warp of the owning 128-thread group
        for (int s = 0; s < PST; ++s) {
            mbar_init(pk_bar(s), 1);                      // packed-ring producer (+tx)
            mbar_init(pk_empty_bar(s), Cfg::NDW);         // decode warps, once the bytes are in registers
        }
        for (int a = 0; a < ACC_STAGES; ++a) {
            mbar_init(tfull_bar(a), 1);                   // tcgen05.commit
            for (int mt = 0; mt < MT; ++mt) mbar_init(tempty_bar(a, mt), CG * 4);  // one arrive per epilogue warp, both CTAs
        }
        fence_mbar_init();
        StallSink* skw = const_cast<StallSink*>(sk);
        skw->buf = p.stall_buf; skw->bar_base = bar_base; skw->nbars = NBARS; skw->cfg = Cfg::ID;
        skw->geom[0] = static_cast<uint32_t>(p.M); skw->geom[1] = static_cast<uint32_t>(p.N);
        skw->geom[2] = static_cast<uint32_t>(p.kb_main); skw->geom[3] = static_cast<uint32_t>(p.kb_tail);
        skw->geom[4] = static_cast<uint32_t>(p.splits); skw->geom[5] = static_cast<uint32_t>(num_tiles);
    }
    if (warp == Cfg::W_ALLOC) {
        tmem_alloc<CG>(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish<CG>();
    }
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may run
    // while the previous kernel of the stream is still draining; nothing below touches global memory before the
    // previous grid has completed and flushed.  Dependents of THIS grid may be scheduled as soon as SMs free up.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if constexpr (Cfg::B_DEC) {
        if (p.am.absmax_q != nullptr)
            for (int i = threadIdx.x; i < 256; i += Cfg::THREADS) code256_s[i] = __ldg(p.am.code256 + i);
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    // barrier addresses as seen for an arrive: the leader's copy when working as a pair
    auto full_bar_arrive = [&](int s) { return CG == 2 ? mapa(full_bar(s), 0) : full_bar(s); };
    auto tempty_bar_arrive = [&](int a, int mt) { return CG == 2 ? mapa(tempty_bar(a, mt), 0) : tempty_bar(a, mt); };

    if (warp == Cfg::W_PROD) {
        // ===================================================== TMA producer ====
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
                int mt_i, nt_i, split;
                tile_coords(p, tile, mt_i, nt_i, split);
                const int m0 = mt_i * p.tile_m;
                const int n0 = nt_i * BN;
                const int kb0 = split * p.kb_main;
                for (int kb = 0; kb < kb_total; ++kb) {
                    const bool tail = kb >= p.kb_main;
                    const int k0 = tail ? (kb - p.kb_main) * 64 : (kb0 + kb) * 64;
                    const CUtensorMap* mapA = tail ? &p.tmA2 : &p.tmA;
                    mbar_wait(empty_bar(s), ph ^ 1u, sk, 1, tile, kb);
                    const bool b_by_tma = tail || !Cfg::B_DEC;
                    const uint32_t tx = Cfg::A_BYTES + (b_by_tma ? Cfg::B_BYTES : 0);
                    const uint32_t fb = full_bar_arrive(s);
                    if constexpr (CG == 2) mbar_arrive_expect_tx_xcta(fb, tx); else mbar_arrive_expect_tx(fb, tx);
                    auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1) {
                        if constexpr (CG == 2) tma_load_2d_pair(dst, map, fb, c0, c1); else tma_load_2d(dst, map, fb, c0, c1);
                    };
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const int row0 = m0 + (CG == 2 ? mt * 256 + static_cast<int>(rank) * 128 : mt * 128);
                        const uint32_t dst = a_stage(s) + mt * 16384;
                        if constexpr (!Cfg::A_MN) {
                            load(dst, mapA, k0, row0);                 // box 64 k x 128 rows
                        } else {
                            load(dst, mapA, row0, k0);                 // box 64 mn x 64 k, two MN chunks
                            load(dst + 8192, mapA, row0 + 64, k0);
                        }
                    }
                    if constexpr (!Cfg::A_MN) {
                        // pull the A tile of k-block kb + pf_dist into L2 now, so that its TMA load is an L2 hit
                        if (p.pf_dist > 0 && !tail && kb + p.pf_dist < p.kb_main) {
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt)
                                tma_prefetch_2d(mapA, k0 + p.pf_dist * 64,
                                                m0 + (CG == 2 ? mt * 256 + static_cast<int>(rank) * 128 : mt * 128));
                        }
                    }
                    const int brow0 = n0 + static_cast<int>(rank) * BNC;
                    if (b_by_tma) {
                        const CUtensorMap* mapB = tail ? &p.tmB2 : &p.tmB;
                        if constexpr (!Cfg::B_MN) {
                            load(b_stage(s), mapB, k0, brow0);         // box 64 k x BNC rows
                        } else {
#pragma unroll
                            for (int c = 0; c < BNC / 64; ++c) load(b_stage(s) + c * 8192, mapB, brow0 + c * 64, k0);
                        }
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
            }
            // drain: every tcgen05.commit aimed at this CTA's empty barriers has landed before exit
            for (int i = 0; i < STAGES; ++i) {
                mbar_wait(empty_bar(s), ph ^ 1u, sk, 2, -1, i);
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == Cfg::W_MMA) {
        // ====================================================== UMMA issuer ====
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(128 * CG, BN, Cfg::A_MN ? 1 : 0, Cfg::B_MN ? 1 : 0);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            int tseq = 0;
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tseq) {
                long long* tr = (p.trace != nullptr && tseq < p.trace_tiles)
                                    ? p.trace + (static_cast<long long>(blockIdx.x) * p.trace_tiles + tseq) * 8 : nullptr;
                long long full_wait = 0;
                for (int kb = 0; kb < kb_total; ++kb) {
                    long long w0 = 0;
                    if (tr != nullptr) {
                        w0 = clock64();
                        if (kb == 0) tr[0] = w0;
                    }
                    mbar_wait<CG == 2>(Cfg::A_XF ? xf_bar(s) : full_bar(s), ph, sk, 3, tile, kb);
                    tc_fence_after();
                    if (tr != nullptr && kb >= STAGES) full_wait += clock64() - w0;
                    const uint64_t a_base = Cfg::A_MN ? umma_desc_sw128(a_stage(s), 8192, 1024) : umma_desc_sw128(a_stage(s), 16, 1024);
                    const uint64_t b_base = Cfg::B_MN ? umma_desc_sw128(b_stage(s), 8192, 1024) : umma_desc_sw128(b_stage(s), 16, 1024);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        if (kb == 0) {   // this sub-tile's accumulator has been drained by the epilogue
                            if (tr != nullptr && mt == 0) tr[1] = clock64();
                            mbar_wait<CG == 2>(tempty_bar(as, mt), aph ^ 1u, sk, 4, tile, mt);
                            tc_fence_after();
                            if (tr != nullptr) tr[2 + (mt > 0)] = clock64();
                        }
                        const uint32_t d_tmem = tmem_base + as * ACC_COLS + mt * BN;
                        // descriptors of (stage, mt, k) = descriptor of the stage base + a constant in the address field
                        // (the stage bases are 1024-byte aligned and the offsets stay far below the 14-bit field)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adesc = a_base + static_cast<uint64_t>((mt * 16384 + k * (Cfg::A_MN ? 2048 : 32)) >> 4);
                            const uint64_t bdesc = b_base + static_cast<uint64_t>((k * (Cfg::B_MN ? 2048 : 32)) >> 4);
                            umma_ss<CG>(d_tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit<CG>(empty_bar(s));            // stage free once these MMAs retire
                    if (kb == kb_total - 1) {
                        umma_commit<CG>(tfull_bar(as));
                        if (tr != nullptr) { tr[4] = clock64(); tr[7] = full_wait; }
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
                if (++as == ACC_STAGES) { as = 0; aph ^= 1u; }
            }
        }
    } else if (Cfg::B_DEC && warp == Cfg::W_PACK) {
        // ================================== TMA producer, packed-NF4 ring ====
        if (lane == 0) {
            int ps = 0;
            uint32_t pph = 0;
            for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
                int mt_i, nt_i, split;
                tile_coords(p, tile, mt_i, nt_i, split);
                const int brow0 = nt_i * BN + static_cast<int>(rank) * BNC;
                const int kb0 = split * p.kb_main;
                for (int kb = 0; kb < p.kb_main; ++kb) {
                    const int k0 = (kb0 + kb) * 64;
                    mbar_wait(pk_empty_bar(ps), pph ^ 1u, sk, 5, tile, kb);
                    mbar_arrive_expect_tx(pk_bar(ps), Cfg::P_BYTES);
                    if constexpr (!Cfg::B_MN)
                        tma_load_2d(p_stage(ps), &p.tmB, pk_bar(ps), k0 / 2, brow0);   // box 32 B x BNC rows of W
                    else
                        tma_load_2d(p_stage(ps), &p.tmB, pk_bar(ps), brow0 / 2, k0);   // box BNC/2 B x 64 rows of W
                    if (++ps == PST) { ps = 0; pph ^= 1u; }
                }
            }
        }
    } else if (warp >= 4 && warp < 4 + 4 * Cfg::ESETS) {
        // ========================================================= epilogue ====
        const int wq = warp & 3;  // TMEM lane quadrant this warp may access
        const int eset = (warp - 4) >> 2;   // epilogue set (0 unless ESETS > 1)
        int as = Cfg::ESETS > 1 ? eset : 0;
        uint32_t aph = 0;
        [[maybe_unused]] const uint32_t stg_base = smem_base + Cfg::RING_BYTES + Cfg::BAR_BYTES + Cfg::CODE256_BYTES;
        int tseq = Cfg::ESETS > 1 ? eset : 0;
        for (int tile = pair_id + (Cfg::ESETS > 1 ? eset * num_pairs : 0); tile < num_tiles; tile += Cfg::ESETS * num_pairs, tseq += Cfg::ESETS) {
            int mt_i, nt_i, split;
            tile_coords(p, tile, mt_i, nt_i, split);
            const int m0 = mt_i * p.tile_m;
            const int n0 = nt_i * BN;
            long long* tr = (p.trace != nullptr && tseq < p.trace_tiles && wq == 0 && lane == 0)
                                ? p.trace + (static_cast<long long>(blockIdx.x) * p.trace_tiles + tseq) * 8 : nullptr;
            mbar_wait(tfull_bar(as), aph, sk, 6, tile, 0);
            tc_fence_after();
            if (tr != nullptr) tr[5] = clock64();
            if constexpr (Cfg::EPI_COAL || Cfg::EPI_TMA) {
                // group q = (mt, c): 32 TMEM lanes (rows) x 64 columns -> bf16 -> swizzled 32 x 128 B staging tile of this
                // warp (row-per-lane writes), then either the copy engine stores / reduce-adds the tile (EPI_TMA) or the
                // warp reads it back 4 rows x 128 B per instruction and writes global memory itself (EPI_COAL).
                constexpr int GPM = BN / 64, NG64 = MT * GPM;
                const uint32_t t_warp = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * ACC_COLS;
                const uint32_t stg = stg_base + (eset * 4 + wq) * 4096;
                const float alpha_r = p.alpha;
                const bool unit = alpha_r == 1.0f;
                __nv_bfloat16* Dp = reinterpret_cast<__nv_bfloat16*>(p.D);
                const long long ldd_r = p.ldd;
                const int M_r = p.M;
                const bool accum_r = p.accum_d != 0;
#pragma unroll 1
                for (int q = 0; q < NG64; ++q) {
                    const int mt = q / GPM, c = q % GPM;
                    const int row_t = m0 + (CG == 2 ? mt * 256 + static_cast<int>(rank) * 128 : mt * 128) + wq * 32;
                    uint32_t o[32];
                    {
                        uint32_t v[32], w[32];
                        tmem_ld_32x32(t_warp + q * 64, v);
                        tmem_ld_32x32(t_warp + q * 64 + 32, w);
                        tmem_ld_wait();
                        if (unit) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                o[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                                o[16 + j] = pack_bf16x2(__uint_as_float(w[2 * j]), __uint_as_float(w[2 * j + 1]));
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                o[j] = pack_bf16x2(__uint_as_float(v[2 * j]) * alpha_r, __uint_as_float(v[2 * j + 1]) * alpha_r);
                                o[16 + j] = pack_bf16x2(__uint_as_float(w[2 * j]) * alpha_r, __uint_as_float(w[2 * j + 1]) * alpha_r);
                            }
                        }
                    }
                    if constexpr (Cfg::EPI == EPI_BF16_MASK) {
                        // LoRA-dropout backward: zero the dropped elements (element index = row * xf_ld + col)
                        const unsigned long long e0 = static_cast<unsigned long long>(row_t + lane) * p.xf_ld + (n0 + c * 64);
                        const uint32_t j0 = static_cast<uint32_t>(e0 >> 2);
                        const uint32_t s_lo = static_cast<uint32_t>(p.seed), s_hi = static_cast<uint32_t>(p.seed >> 32);
                        const uint32_t thr = p.thresh16, flip = p.mask_flip;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {   // one 64-bit hash per four elements (two packed words)
                            uint32_t ha, hb;
                            dropout_hash64(s_lo, s_hi, j0 + j, ha, hb);
                            o[2 * j] &= dropout_mask2(ha, thr) ^ flip;
                            o[2 * j + 1] &= dropout_mask2(hb, thr) ^ flip;
                        }
                    }
                    if (c == GPM - 1) {   // sub-tile mt drained: its last TMEM load has completed
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (CG == 2) mbar_arrive_xcta(tempty_bar_arrive(as, mt)); else mbar_arrive(tempty_bar(as, mt));
                        }
                    }
                    if constexpr (Cfg::EPI_TMA) {
                        if (lane == 0) tma_store_wait_read<0>();   // the copy engine has read the previous group
                    }
                    __syncwarp();   // (EPI_COAL) the previous group's read-back of the staging tile is complete
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg + lane * 128 + ((static_cast<uint32_t>(j ^ (lane & 7))) << 4)),
                                     "r"(o[4 * j]), "r"(o[4 * j + 1]), "r"(o[4 * j + 2]), "r"(o[4 * j + 3]) : "memory");
                    if constexpr (Cfg::EPI_TMA) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (row_t < M_r) {
                                if (accum_r) tma_reduce_add_2d(&p.tmD, stg, n0 + c * 64, row_t);
                                else tma_store_2d(&p.tmD, stg, n0 + c * 64, row_t);
                            }
                            tma_store_commit();
                        }
                    } else {
                        __syncwarp();
                        const int row_w = row_t + (lane >> 3);
                        const int ch = lane & 7;
                        uint4 val[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {   // rows i*4 + lane/8
                            const int r = i * 4 + (lane >> 3);
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(val[i].x), "=r"(val[i].y), "=r"(val[i].z), "=r"(val[i].w)
                                         : "r"(stg + r * 128 + ((static_cast<uint32_t>(ch ^ (r & 7))) << 4)));
                        }
                        __nv_bfloat16* gp = Dp + static_cast<long long>(row_w) * ldd_r + (n0 + c * 64 + ch * 8);
                        if (accum_r) {
                            // D += tile: 128-bit vector reductions performed at the L2 (fire and forget, no read into
                            // the SM); bf16 add with round-to-nearest, i.e. D = bf16(D + bf16(alpha * acc)).  A vector of
                            // eight zeros changes nothing and is not sent: the dX correction keeps only the ~5 % of
                            // elements the dropout mask dropped, so two thirds of its vectors are empty and the L2
                            // reduction rate, which bounds this kernel, is spent on the third that is not.
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t ok = (row_w + i * 4 < M_r && (val[i].x | val[i].y | val[i].z | val[i].w) != 0u) ? 1u : 0u;
                                asm volatile("{\n\t.reg .pred P1;\n\tsetp.ne.b32 P1, %5, 0;\n\t@P1 red.global.add.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};\n\t}\n"
                                             ::"l"(gp + static_cast<long long>(i) * 4 * ldd_r), "r"(val[i].x), "r"(val[i].y),
                                               "r"(val[i].z), "r"(val[i].w), "r"(ok) : "memory");
                            }
                        } else if (p.stream_out) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t ok = (row_w + i * 4 < M_r) ? 1u : 0u;
                                asm volatile("{\n\t.reg .pred P1;\n\tsetp.ne.b32 P1, %5, 0;\n\t@P1 st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};\n\t}\n"
                                             ::"l"(gp + static_cast<long long>(i) * 4 * ldd_r), "r"(val[i].x), "r"(val[i].y),
                                               "r"(val[i].z), "r"(val[i].w), "r"(ok) : "memory");
                            }
                        } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t ok = (row_w + i * 4 < M_r) ? 1u : 0u;
                            asm volatile("{\n\t.reg .pred P1;\n\tsetp.ne.b32 P1, %5, 0;\n\t@P1 st.global.v4.b32 [%0], {%1,%2,%3,%4};\n\t}\n"
                                         ::"l"(gp + static_cast<long long>(i) * 4 * ldd_r), "r"(val[i].x), "r"(val[i].y),
                                           "r"(val[i].z), "r"(val[i].w), "r"(ok) : "memory");
                        }
                        }
                    }
                }
            } else
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int row = m0 + (CG == 2 ? mt * 256 + static_cast<int>(rank) * 128 : mt * 128) + wq * 32 + lane;
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * ACC_COLS + mt * BN;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    const int col0 = n0 + c * 32;
                    if (row < p.M && row - m0 < p.tile_m && col0 < p.N) {
                        if constexpr (Cfg::EPI == EPI_BF16 || Cfg::EPI == EPI_BF16_MASK) {
                            uint32_t o[16];
                            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.D) +
                                                                  static_cast<long long>(row) * p.ldd + col0);
                            if constexpr (Cfg::EPI == EPI_BF16_MASK) {
                                const unsigned long long e0 = static_cast<unsigned long long>(row) * p.xf_ld + col0;
                                const uint32_t j0 = static_cast<uint32_t>(e0 >> 2);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    uint32_t ha, hb;
                                    dropout_hash64(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32), j0 + j, ha, hb);
                                    o[2 * j] = pack_bf16x2(__uint_as_float(v[4 * j]) * p.alpha, __uint_as_float(v[4 * j + 1]) * p.alpha) &
                                               (dropout_mask2(ha, p.thresh16) ^ p.mask_flip);
                                    o[2 * j + 1] = pack_bf16x2(__uint_as_float(v[4 * j + 2]) * p.alpha, __uint_as_float(v[4 * j + 3]) * p.alpha) &
                                                   (dropout_mask2(hb, p.thresh16) ^ p.mask_flip);
                                }
                            } else if (p.accum_d) {
#pragma unroll
                                for (int j4 = 0; j4 < 4; ++j4) {
                                    const uint4 old = dst[j4];
                                    const uint32_t ow[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
                                    for (int jj = 0; jj < 4; ++jj) {
                                        const int j = 4 * j4 + jj;
                                        const __nv_bfloat162 ob = *reinterpret_cast<const __nv_bfloat162*>(&ow[jj]);
                                        o[j] = pack_bf16x2(__bfloat162float(ob.x) + __uint_as_float(v[2 * j]) * p.alpha,
                                                           __bfloat162float(ob.y) + __uint_as_float(v[2 * j + 1]) * p.alpha);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    o[j] = pack_bf16x2(__uint_as_float(v[2 * j]) * p.alpha,
                                                       __uint_as_float(v[2 * j + 1]) * p.alpha);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                            if (Cfg::EPI == EPI_BF16 && p.D2 != nullptr) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    o[j] = pack_bf16x2(__uint_as_float(v[2 * j]) * p.alpha2,
                                                       __uint_as_float(v[2 * j + 1]) * p.alpha2);
                                uint4* dst2 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.D2) +
                                                                       static_cast<long long>(row) * p.ldd + col0);
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    dst2[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                            }
                        } else if constexpr (Cfg::EPI == EPI_F32_PART) {
                            float* dst = reinterpret_cast<float*>(p.D) +
                                         (static_cast<long long>(split) * p.M + row) * p.N + col0;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                reinterpret_cast<uint4*>(dst)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        } else {
                            float* dst = reinterpret_cast<float*>(p.D) +
                                         (static_cast<long long>(split) * p.N + col0) * p.M + row;
#pragma unroll
                            for (int j = 0; j < 32; ++j) dst[static_cast<long long>(j) * p.M] = __uint_as_float(v[j]);
                        }
                    }
                }
                tc_fence_before();   // sub-tile mt drained
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) mbar_arrive_xcta(tempty_bar_arrive(as, mt)); else mbar_arrive(tempty_bar(as, mt));
                }
            }
            if (tr != nullptr) tr[6] = clock64();
            if constexpr (Cfg::ESETS > 1) { aph ^= 1u; } else { if (++as == ACC_STAGES) { as = 0; aph ^= 1u; } }
        }
        if constexpr (Cfg::EPI_TMA) {
            if (lane == 0) tma_store_wait<0>();   // all bulk stores of this warp are complete before the CTA exits
        }
    } else if (warp >= 8 || Cfg::B_DEC) {
      if constexpr (Cfg::A_XF) {
        // ============================================ A-operand transform ====
        // LoRA dropout: zero the dropped elements of the A tile in place (the 1/(1-p) scale is folded
        // into the epilogue).  XF_GROUPS groups of 128 threads take ring positions in turn, so several stages are
        // being transformed at any time (one group alone is latency-bound: barrier wake-up, LDS, hash, STS,
        // proxy fence, arrive per stage).  A thread handles eight 16-byte chunks (8 bf16 each) of its stage;
        // chunk q sits at byte q*16: row q>>3, physical chunk q&7 holds logical chunk (q&7)^(row&7).
        constexpr int XG = Cfg::XF_GROUPS, XT = Cfg::XF_THREADS / XG;
        const int g = (threadIdx.x - 256) / XT;
        const int t = (threadIdx.x - 256) % XT;
        const uint32_t seed_lo = static_cast<uint32_t>(p.seed), seed_hi = static_cast<uint32_t>(p.seed >> 32);
        const uint32_t thr = p.thresh16;
        uint32_t it = 0;   // ring position over all k-blocks of all tiles
        for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
            int mt_i, nt_i, split;
            tile_coords(p, tile, mt_i, nt_i, split);
            const int row0 = mt_i * p.tile_m;
            const int kb0 = split * p.kb_main;
            for (int kb = static_cast<int>((static_cast<uint32_t>(g) - it) & (XG - 1)); kb < kb_total; kb += XG) {
                const int k0 = (kb0 + kb) * 64;
                const uint32_t si = it + kb, s = si % STAGES, ph = (si / STAGES) & 1u;
                // Consecutive uses of a stage are transformed by DIFFERENT groups whenever XG does not divide STAGES, and
                // a parity wait cannot tell "use u has landed" from "use u-1 has not landed yet".  A group that runs
                // ahead (its previous stage landed before an older load did -- TMA completions are not ordered) would
                // transform the stage one use early and complete xf_bar(s) twice before the issuer looked once: the
                // round-1 stall.  Waiting first for the previous use of the stage to be RELEASED
                // (empty_bar phase u-1: exact, because this group's previous position si-XG was issued by the producer
                // after it had seen phase u-2) puts full_bar(s) at phase >= u, which makes the parity wait exact too.
                static_assert(XG <= STAGES, "transform groups: the producer-order argument above needs XG <= STAGES");
                mbar_wait(empty_bar(s), ph ^ 1u, sk, 16, tile, kb);
                mbar_wait(full_bar(s), ph, sk, 8, tile, kb);
                // hash counter of chunk i = counter of chunk 0 + a multiple of xf_ld (XT = 128: rows advance by 16 per chunk)
                uint32_t j0_base;
                {
                    const int r0 = t >> 3, cl = (t & 7) ^ (r0 & 7);
                    const long long e_base = Cfg::A_MN ? static_cast<long long>(k0 + r0) * p.xf_ld + row0 + cl * 8
                                                       : static_cast<long long>(row0 + r0) * p.xf_ld + k0 + cl * 8;
                    j0_base = static_cast<uint32_t>(static_cast<unsigned long long>(e_base) >> 2);
                }
                const uint32_t j0_step = static_cast<uint32_t>(p.xf_ld) * 4u;   // 16 rows further = 16 * xf_ld elements = 4 * xf_ld counters
#pragma unroll
                for (int i = 0; i < 1024 / XT; ++i) {
                    const int q = t + XT * i;
                    static_assert(XT == 128, "chunk -> row mapping below assumes 128 threads per transform group");
                    const uint32_t j0 = Cfg::A_MN ? j0_base + static_cast<uint32_t>(i & 3) * j0_step + static_cast<uint32_t>(i >> 2) * 16u
                                                  : j0_base + static_cast<uint32_t>(i) * j0_step;
                    const uint32_t addr = a_stage(s) + q * 16;
                    uint32_t w[4];
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
#pragma unroll
                    for (int j = 0; j < 2; ++j) {   // 8 elements = two 64-bit hashes
                        uint32_t ha, hb;
                        dropout_hash64(seed_lo, seed_hi, j0 + j, ha, hb);
                        w[2 * j] &= dropout_mask2(ha, thr);
                        w[2 * j + 1] &= dropout_mask2(hb, thr);
                    }
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                                 "r"(w[3]) : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(xf_bar(s));
            }
            it += kb_total;
        }
      }
      if constexpr (Cfg::B_DEC) {
        // ======================================================= NF4 decode ====
        // NG groups of BNC threads; group g takes ring positions it with it % NG == g.  One thread
        // = one quantisation block (64 weights, one absmax, one 128-byte operand row) per stage.
        // Decode warps are the warp ids {0, 2, 3, 8, 10, 11, 12, 14} (see the role map in GemmCfg); 15 is idle.
        const int wi = warp < 4 ? (warp == 0 ? 0 : warp - 1) : 3 + (warp - 8) - ((warp - 6) >> 2);   // index among them
        if (wi >= NG * Cfg::NDW) goto decode_done;
        {
        const int g = wi / Cfg::NDW;
        const int t = (wi % Cfg::NDW) * 32 + lane;
        float code16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) code16[i] = __ldg(p.code16 + i);
        const bool nested = p.am.absmax_q != nullptr;
        auto absmax_of = [&](long long blk) -> float {
            if (!nested) return __ldg(p.am.absmax_f32 + blk);
            const float c = code256_s[__ldg(p.am.absmax_q + blk)];
            return __fadd_rn(__fmul_rn(c, __ldg(p.am.absmax2 + (blk >> 8))), p.am.offset);
        };
        // destination row of this thread inside the B stage and its swizzle key
        constexpr int PARTS = BNC / 64;  // MN-major: 64-wide chunks of the output dimension
        const int r_loc = Cfg::B_MN ? t / PARTS : t;
        const int part = Cfg::B_MN ? t % PARTS : 0;
        const uint32_t dst_off = Cfg::B_MN ? static_cast<uint32_t>((part * 8 + (r_loc >> 3)) * 1024 + (r_loc & 7) * 128)
                                           : static_cast<uint32_t>((r_loc >> 3) * 1024 + (r_loc & 7) * 128);
        const uint32_t key = static_cast<uint32_t>(r_loc & 7);
        uint32_t it = 0;   // ring position over all k-blocks (main + tail) of all tiles
        uint32_t pit = 0;  // ring position over main k-blocks only (packed ring)
        for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
            int mt_i, nt_i, split;
            tile_coords(p, tile, mt_i, nt_i, split);
            const int n0 = nt_i * BN + static_cast<int>(rank) * BNC;
            // absmax block index of k-block kb:  K-major: (n0 + t) * kpr + kb
            //                                    MN-major: (kb*64 + r_loc) * kpr + (n0 + part*64)/64
            const long long blk_step = Cfg::B_MN ? 64LL * p.kpr : 1LL;
            const long long blk0 = (Cfg::B_MN ? static_cast<long long>(r_loc) * p.kpr + (n0 >> 6) + part
                                              : static_cast<long long>(n0 + t) * p.kpr) +
                                   static_cast<long long>(split) * p.kb_main * blk_step;
            // first main k-block of this tile that belongs to this group
            int kb_first = static_cast<int>((static_cast<uint32_t>(g) - it) & (NG - 1));
            float a_next = kb_first < p.kb_main ? absmax_of(blk0 + kb_first * blk_step) : 0.f;
            for (int kb = kb_first; kb < p.kb_main; kb += NG) {
                const float a = a_next;
                if (kb + NG < p.kb_main) a_next = absmax_of(blk0 + (kb + NG) * blk_step);  // prefetch
                Nf4Lut lut;
                nf4_build_lut(code16, a, lut);
                const uint32_t pi = pit + kb, ps = pi % PST, pph = (pi / PST) & 1u;
                // The group that reads packed slot ps alternates at tile boundaries when kb_tail is odd (`it` and `pit`
                // drift apart by one per tile), so this group may not have seen the slot's previous use: wait until that
                // use has been released by its readers (pk_empty_bar phase u-1 -- exact: this group's previous position
                // pi-1 or pi-2 was issued by the producer after it had seen phase u-2), after which pk_bar(ps) is at
                // phase >= u and the parity wait below cannot alias "use u-1 not landed yet" with "use u landed".
                static_assert(NG <= PST, "decode groups: the producer-order argument above needs NG <= PST");
                mbar_wait(pk_empty_bar(ps), pph ^ 1u, sk, 17, tile, kb);
                mbar_wait(pk_bar(ps), pph, sk, 9, tile, kb);
                uint32_t w[8];
                {
                    const uint32_t src = p_stage(ps) + t * 32;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(src));
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(src + 16));
                }
                {
                    // Release the packed slot only after every lane's loads have RETURNED.  Issuing the arrive right behind
                    // the two LDS is not enough: the shared-memory pipe of these kernels is ~75 % busy (operand reads of the
                    // tensor core, TMA writes, 32 KB of decoded stores per k-block), a load can sit in its queue for longer
                    // than the refill of the slot takes, and the TMA then overwrites the first rows of the slot under a
                    // reader that has not read yet: one decode warp's 32 weight rows of one k-block come from the NEXT use
                    // of the slot -- a few tiles per thousand launches were off by 3-14 % of max|y| (round 2, DESIGN.md
                    // section 4; whether it happened depended on where ptxas scheduled the first consumer of w[]).  The
                    // store to a scratch word consumes the loaded registers and keeps the dependency alive.
#ifdef B2Q_FENCE_REDUX
                    // first form of the fix (validated, 1 % slower): a warp-wide reduction reads all eight words of all lanes
                    const uint32_t x = __reduce_xor_sync(0xffffffffu, w[0] ^ w[1] ^ w[2] ^ w[3] ^ w[4] ^ w[5] ^ w[6] ^ w[7]);
#else
                    // one consumer of the last word of each LDS.128: the scoreboard of a load covers the whole warp
                    // instruction, so the store below cannot issue before both loads have returned for all 32 lanes
                    // (validated like the first form: 0 bad launches in tools/dx_check.py / kb_probe.py, +0.5 % step)
                    const uint32_t x = w[3] ^ w[7];
                    __syncwarp();
#endif
                    if (lane == 0) {
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(tmem_slot + 8u), "r"(x) : "memory");
                        mbar_arrive(pk_empty_bar(ps));  // packed slot may be refilled
                    }
                }
                const uint32_t si = it + kb, s = si % STAGES, ph = (si / STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u, sk, 10, tile, kb);   // decoded-B slot free (MMAs that read it retired)
                const uint32_t dst = b_stage(s) + dst_off;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t o[4];
                    nf4_decode_word(w[j], lut, o);
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst + ((static_cast<uint32_t>(j) ^ key) << 4)),
                                 "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) mbar_arrive_xcta(full_bar_arrive(s)); else mbar_arrive(full_bar(s));
                }
            }
            // LoRA tail k-blocks: B arrives by TMA; the group that owns the ring position keeps the count
            for (int kb = p.kb_main; kb < kb_total; ++kb) {
                const uint32_t si = it + kb;
                if ((si & (NG - 1)) != static_cast<uint32_t>(g)) continue;
                const uint32_t s = si % STAGES, ph = (si / STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u, sk, 11, tile, kb);
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) mbar_arrive_xcta(full_bar_arrive(s)); else mbar_arrive(full_bar(s));
                }
            }
            it += kb_total;
            pit += p.kb_main;
        }
        }
        decode_done:;
      }
    }

    // ------------------------------------------------------------- teardown ----
    __syncwarp();  // single-lane roles (producer, UMMA issuer) rejoin their warp before the aligned barrier
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync(); else __syncthreads();
    if (warp == Cfg::W_ALLOC) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace b2q
