// In-register NF4 decode shared by the stand-alone decode kernel (algo 1) and the
// tcgen05 GEMM main loops.  Bit-exact restatement of bitsandbytes' dequantize_4bit
// (SURVEY.md section 8a row a7):  w = bf16_rn(fl32(code16[nibble] * absmax)).
//
// Scheme ("pre-scaled LUT + PRMT"): one thread owns one quantisation block (64 weights,
// 32 packed bytes, ONE absmax).  It builds the 16 possible results once
// (16 FMUL + 8 cvt.rn.bf16x2.f32 - the very same fp32 product and RN conversion the
// reference performs per element), splits them into low-byte / high-byte planes
// (4 + 4 registers), and then looks 4 nibbles up at a time with PRMT, whose selector
// nibbles are the packed nibbles themselves (&7), plus a sign-replicating PRMT that
// turns nibble bit 3 into a byte mask to choose between the two table halves.
// Cost ~2.9 ALU-pipe ops per weight and no shared-memory traffic.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

namespace b2q {

// Raw PTX prmt (default mode).  NOT __byte_perm: that intrinsic ANDs the selector with 0x7777,
// which would strip the sign-replicate bit (selector nibble bit 3) the mask generation relies on.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

struct Nf4Lut {
    uint32_t L[4];  // low bytes of entries 0..15  (entry e in byte e%4 of L[e/4])
    uint32_t H[4];  // high bytes
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x = lo (low half), .y = hi
    return *reinterpret_cast<uint32_t*>(&v);
}

// code16: the 16 fp32 code values (QuantState.code / "quant_map"), absmax: fp32 scale of the block.
__device__ __forceinline__ void nf4_build_lut(const float (&code16)[16], float absmax, Nf4Lut& lut) {
    uint32_t P[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        P[j] = pack_bf16x2(__fmul_rn(code16[2 * j], absmax), __fmul_rn(code16[2 * j + 1], absmax));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        lut.L[j] = prmt(P[2 * j], P[2 * j + 1], 0x6420);
        lut.H[j] = prmt(P[2 * j], P[2 * j + 1], 0x7531);
    }
}

// Decode one packed 32-bit word (4 bytes = 8 weights e0..e7, byte i = e(2i)<<4 | e(2i+1))
// into 4 registers of bf16x2 in element order (out[0] = e0 | e1<<16, ...).
__device__ __forceinline__ void nf4_decode_word(uint32_t w, const Nf4Lut& lut, uint32_t (&out)[4]) {
    const uint32_t w7 = w & 0x77777777u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        // selector nibbles p0..p3 of this half hold elements (e1, e0, e3, e2)
        const uint32_t s7 = h ? (w7 >> 16) : w7;
        const uint32_t x = h ? (w >> 16) : w;
        const uint32_t y = h ? (w >> 12) : (w << 4);
        const uint32_t la = prmt(lut.L[0], lut.L[1], s7);
        const uint32_t lb = prmt(lut.L[2], lut.L[3], s7);
        const uint32_t ha = prmt(lut.H[0], lut.H[1], s7);
        const uint32_t hb = prmt(lut.H[2], lut.H[3], s7);
        // byte i of m = 0xFF iff bit 3 of nibble p_i is set (sign-replicate mode of PRMT)
        const uint32_t m = prmt(x, y, 0x9D8C);
        const uint32_t lo = (la & ~m) | (lb & m);
        const uint32_t hi = (ha & ~m) | (hb & m);
        out[2 * h + 0] = prmt(lo, hi, 0x4051);
        out[2 * h + 1] = prmt(lo, hi, 0x6273);
    }
}

// absmax of one 64-block.  Plain: fp32 vector.  Nested (double quant):
//   fl32(fl32(code256[q] * absmax2[blk / 256]) + offset)   -- no FMA contraction.
struct AbsmaxSrc {
    const float* absmax_f32;     // plain, or nullptr
    const uint8_t* absmax_q;     // nested
    const float* absmax2;        // nested, one per 256 blocks
    const float* code256;        // nested, 256 entries
    float offset;                // nested
};

__device__ __forceinline__ float load_absmax(const AbsmaxSrc& s, long long blk) {
    if (s.absmax_q == nullptr) return __ldg(s.absmax_f32 + blk);
    const float c = __ldg(s.code256 + __ldg(s.absmax_q + blk));
    const float a2 = __ldg(s.absmax2 + (blk >> 8));
    return __fadd_rn(__fmul_rn(c, a2), s.offset);
}

}  // namespace b2q
