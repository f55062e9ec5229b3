/* Plain-C restatement of the NF4 blockwise quantise / dequantise arithmetic.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.
 *
 * Follows the published bitsandbytes algorithm (kQuantizeBlockwise / kDequantizeBlockwise with
 * DATA_TYPE = NF4), which the reference reaches through
 * cullavo/load_cullavo.py:73-86.  Second, independent implementation next to oracle/nf4.py:
 * tests/test_oracle_c.py checks that the two agree bit for bit.
 *
 * gcc -O2 -shared -fPIC -o oracle/_build/libnf4ref.so oracle/nf4_ref.c   (no -ffast-math: fp32 op order matters)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const float NF4_THRESH[15] = {
    -0.8480964004993439f,  -0.6106329262256622f,  -0.4599952697753906f, -0.33967943489551544f, -0.23460740596055984f,
    -0.13791173323988914f, -0.045525018125772476f, 0.03979014977812767f, 0.1202552504837513f,  0.2035212516784668f,
    0.2920137718319893f,   0.3893125355243683f,   0.5016634166240692f,  0.6427869200706482f,  0.8614784181118011f};

/* the comparison tree of dQuantizeNF4, written out as a binary search with strict '>' */
static uint8_t nf4_code(float x) {
    int lo = 0, hi = 15; /* answer = number of thresholds t with x > t */
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (x > NF4_THRESH[mid]) lo = mid + 1; else hi = mid;
    }
    return (uint8_t)lo; /* NaN: every comparison false -> 0 */
}

static uint16_t bf16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

/* n values -> packed[(n+1)/2], absmax[ceil(n/64)] */
void nf4ref_quantize(const float* w, int64_t n, uint8_t* packed, float* absmax) {
    int64_t nblocks = (n + 63) / 64;
    for (int64_t b = 0; b < nblocks; ++b) {
        int64_t i0 = b * 64, i1 = i0 + 64 < n ? i0 + 64 : n;
        float a = 0.0f;
        for (int64_t i = i0; i < i1; ++i) { float v = fabsf(w[i]); if (v > a) a = v; }
        absmax[b] = a;
        volatile float inv = 1.0f / a; /* inf for an all-zero block */
        for (int64_t i = i0; i < i1; i += 2) {
            volatile float x0 = w[i] * inv;
            uint8_t c0 = nf4_code(x0), c1 = 0;
            if (i + 1 < i1) { volatile float x1 = w[i + 1] * inv; c1 = nf4_code(x1); }
            packed[i >> 1] = (uint8_t)((c0 << 4) | c1);
        }
    }
}

/* absmax_q == NULL: plain fp32 absmax; else nested: fl32(fl32(code256[q] * absmax2[j/256]) + offset) */
void nf4ref_dequantize_bf16(const uint8_t* packed, const float* absmax, const uint8_t* absmax_q, const float* absmax2,
                            const float* code256, float offset, const float* code16, int64_t n, uint16_t* out) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t b = i / 64;
        float a;
        if (absmax_q) {
            volatile float prod = code256[absmax_q[b]] * absmax2[b / 256];
            a = prod + offset;
        } else {
            a = absmax[b];
        }
        uint8_t byte = packed[i >> 1];
        uint8_t nib = (i & 1) ? (byte & 0x0f) : (byte >> 4);
        volatile float v = code16[nib] * a;
        out[i] = bf16_rn(v);
    }
}
