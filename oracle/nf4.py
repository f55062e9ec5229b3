"""NF4 blockwise quantise / dequantise, restated on the CPU with numpy (fp32 op order kept).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.

What this restates (third-party, absent from /root/reference; reached from the
reference only through cullavo/load_cullavo.py:73-82,86):

* bitsandbytes ``functional.quantize_4bit`` -> CUDA ``kQuantizeBlockwise<T,64,2,0,NF4>``
  (SURVEY.md section 8a row a6): per 64-element block ``a = max|w|``; the code of an
  element is the number of the 15 NF4 mid-point thresholds that are strictly
  below ``w * (1.0f / a)`` (that is what the ``dQuantizeNF4`` comparison tree
  computes); two codes per byte, EVEN element in the HIGH nibble.
* double quantisation of the absmax vector (``compress_statistics=True``):
  ``offset = mean(absmax)``; ``absmax - offset`` is quantised blockwise (256) to the
  nearest entry of the signed 8-bit "dynamic map" with one fp32 scale per block.
* bitsandbytes ``functional.dequantize_4bit`` -> ``kDequantizeBlockwise<bf16,512,64,8,NF4>``
  (row a7):  ``absmax_f32[j] = fl32(fl32(code256[q[j]] * absmax2[j // 256]) + offset)``,
  ``W[i] = bf16_rn(fl32(NF4[nib(i)] * absmax_f32[i // 64]))``.
"""
from __future__ import annotations

import numpy as np
import torch

# The 16 NF4 code values (index = nibble), fp32.  QLoRA construction: quantiles of
# N(0,1), asymmetric, normalised to [-1, 1]  (checked in tests/test_oracle.py).
NF4_CODE = np.array(
    [
        -1.0,
        -0.6961928009986877,
        -0.5250730514526367,
        -0.39491748809814453,
        -0.28444138169288635,
        -0.18477343022823334,
        -0.09105003625154495,
        0.0,
        0.07958029955625534,
        0.16093020141124725,
        0.24611230194568634,
        0.33791524171829224,
        0.44070982933044434,
        0.5626170039176941,
        0.7229568362236023,
        1.0,
    ],
    dtype=np.float32,
)

# The 15 decision thresholds of the quantiser's comparison tree (fp32 literals).
NF4_THRESHOLDS = np.array(
    [
        -0.8480964004993439,
        -0.6106329262256622,
        -0.4599952697753906,
        -0.33967943489551544,
        -0.23460740596055984,
        -0.13791173323988914,
        -0.045525018125772476,
        0.03979014977812767,
        0.1202552504837513,
        0.2035212516784668,
        0.2920137718319893,
        0.3893125355243683,
        0.5016634166240692,
        0.6427869200706482,
        0.8614784181118011,
    ],
    dtype=np.float32,
)


def create_dynamic_map(signed: bool = True, max_exponent_bits: int = 7, total_bits: int = 8) -> np.ndarray:
    """The 256-entry 8-bit "dynamic" code used for the nested absmax.

    Follows bitsandbytes ``functional.create_dynamic_map``: for exponent step ``i`` the
    bin means of ``linspace(0.1, 1, 2**i + 1)`` scaled by ``10**(i - 6)``, with both signs,
    plus 0 and 1.0; sorted ascending.  The linspace / mean arithmetic is fp32 (torch
    default dtype), the power-of-ten scale is a python float, the result is stored fp32.
    """
    data: list[float] = []
    non_sign_bits = total_bits - 1
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    i = 0
    for i in range(max_exponent_bits):
        if signed:
            fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1)
        else:
            fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        boundaries = torch.linspace(0.1, 1, fraction_items)  # fp32, as upstream
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        scale = 10 ** (-(max_exponent_bits - 1) + i)
        data += (scale * means).tolist()
        if signed:
            data += (-(scale) * means).tolist()
    if additional_items > 0:
        boundaries = torch.linspace(0.1, 1, additional_items + 1)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        scale = 10 ** (-(max_exponent_bits - 1) + i)
        data += (scale * means).tolist()
        if signed:
            data += (-(scale) * means).tolist()
    data.append(0.0)
    data.append(1.0)
    assert len(data) == 2**total_bits
    data.sort()
    return np.asarray(data, dtype=np.float32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit patterns (uint16), round-to-nearest-even, NaN kept quiet."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    lsb = (u >> np.uint32(16)) & np.uint32(1)
    rounded = (u + np.uint32(0x7FFF) + lsb) >> np.uint32(16)
    nan = np.isnan(x)
    out = rounded.astype(np.uint16)
    if nan.any():
        out = np.where(nan, ((u >> np.uint32(16)) | np.uint32(0x0040)).astype(np.uint16), out)
    return out


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


def _nf4_codes(xn: np.ndarray) -> np.ndarray:
    """Number of thresholds strictly below xn (== the comparison tree); NaN -> 0."""
    codes = np.searchsorted(NF4_THRESHOLDS, xn, side="left").astype(np.uint8)
    codes[np.isnan(xn)] = 0
    return codes


def _nearest_code(code: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Index of the nearest entry of the sorted ``code``; ties go to the lower index."""
    hi = np.clip(np.searchsorted(code, x, side="left"), 1, len(code) - 1)
    lo = hi - 1
    pick_hi = (code[hi] - x) < (x - code[lo])
    return np.where(pick_hi, hi, lo).astype(np.uint8)


def quantize_nf4(w: np.ndarray, blocksize: int = 64, double_quant: bool = False) -> dict:
    """Quantise a weight (any shape, row-major order) to packed NF4.

    Returns a dict mirroring bitsandbytes' ``QuantState`` fields:
      packed   uint8 [(n+1)//2]      two codes per byte, even element in the high nibble
      absmax   fp32 [nblocks]        (plain)            -- or --
      absmax_q uint8 [nblocks], absmax2 fp32 [ceil(nblocks/256)], code256 fp32 [256],
      offset   fp32 scalar           (double_quant)
      code     fp32 [16], shape, blocksize
    """
    shape = tuple(w.shape)
    flat = np.ascontiguousarray(w, dtype=np.float32).reshape(-1)
    n = flat.size
    nblocks = (n + blocksize - 1) // blocksize
    pad = nblocks * blocksize - n
    blk = np.pad(flat, (0, pad)).reshape(nblocks, blocksize)
    absmax = np.abs(blk).max(axis=1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (np.float32(1.0) / absmax).astype(np.float32)  # 1.0f / absmax, inf for a zero block
        xn = (blk * inv[:, None]).astype(np.float32)  # 0 * inf = NaN -> code 0
    codes = _nf4_codes(xn).reshape(-1)[:n]
    if n % 2:
        codes = np.concatenate([codes, np.zeros(1, np.uint8)])
    packed = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)
    state = {
        "packed": packed,
        "code": NF4_CODE.copy(),
        "shape": shape,
        "blocksize": blocksize,
        "nested": bool(double_quant),
    }
    if not double_quant:
        state["absmax"] = absmax
        return state
    state.update(double_quantize_absmax(absmax))
    return state


def double_quantize_absmax(absmax: np.ndarray) -> dict:
    """Second-level quantisation of the fp32 absmax vector (bitsandbytes ``quantize_4bit(compress_statistics=True)``):
    ``offset = absmax.mean(); absmax -= offset; quantize_blockwise(absmax, blocksize=256)`` with the 8-bit dynamic map.

    The mean: bitsandbytes takes ``absmax.mean()`` with torch's CUDA reduction (fp32 partial sums in a device- and
    version-dependent tree order, so not reproducible bit for bit by any other implementation); this restatement and the
    product's ``absmax_mean_kernel`` return the CORRECTLY ROUNDED mean (fp64 accumulate, one rounding), which differs from
    any fp32 summation order by at most an ulp or two.  The offset is part of the serialised quant state, so decode
    parity does not depend on it; quantiser BYTES may differ from bitsandbytes' own in the rare absmax value that sits
    within that distance of an 8-bit code boundary (tests/test_oracle.py bounds the effect on the decoded weight)."""
    absmax = np.ascontiguousarray(absmax, dtype=np.float32)
    nblocks = absmax.size
    offset = np.float32(absmax.astype(np.float64).mean())
    centred = (absmax - offset).astype(np.float32)
    code256 = create_dynamic_map()
    nb2 = (nblocks + 255) // 256
    blk2 = np.pad(centred, (0, nb2 * 256 - nblocks)).reshape(nb2, 256)
    absmax2 = np.abs(blk2).max(axis=1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv2 = (np.float32(1.0) / absmax2).astype(np.float32)
        xn2 = (blk2 * inv2[:, None]).astype(np.float32)
    xn2 = np.nan_to_num(xn2, nan=0.0)
    q = _nearest_code(code256, xn2).reshape(-1)[:nblocks]
    return dict(absmax_q=q, absmax2=absmax2, code256=code256, offset=offset)


def dequantize_absmax(state: dict) -> np.ndarray:
    """fp32 absmax per 64-block; nested: ``fl32(fl32(code256[q] * absmax2[j//256]) + offset)``."""
    if not state["nested"]:
        return state["absmax"].astype(np.float32)
    q = state["absmax_q"]
    j = np.arange(q.size) // 256
    prod = (state["code256"][q].astype(np.float32) * state["absmax2"][j].astype(np.float32)).astype(np.float32)
    return (prod + np.float32(state["offset"])).astype(np.float32)


def unpack_codes(packed: np.ndarray, n: int) -> np.ndarray:
    codes = np.empty(packed.size * 2, np.uint8)
    codes[0::2] = packed >> 4
    codes[1::2] = packed & 0x0F
    return codes[:n]


def dequantize_nf4(state: dict, as_bits: bool = True) -> np.ndarray:
    """Decode to bf16.  Returns uint16 bit patterns (default) or the fp32 values of them."""
    n = int(np.prod(state["shape"]))
    bs = state["blocksize"]
    codes = unpack_codes(state["packed"].reshape(-1), n)
    absmax = dequantize_absmax(state)
    vals = (state["code"].astype(np.float32)[codes] * absmax[np.arange(n) // bs]).astype(np.float32)
    bits = bf16_round(vals).reshape(state["shape"])
    return bits if as_bits else bf16_bits_to_f32(bits)


def dequantize_nf4_f32(state: dict) -> np.ndarray:
    """Decode without the final bf16 rounding (what a fp32 ``compute_dtype`` would see)."""
    n = int(np.prod(state["shape"]))
    codes = unpack_codes(state["packed"].reshape(-1), n)
    absmax = dequantize_absmax(state)
    vals = state["code"].astype(np.float32)[codes] * absmax[np.arange(n) // state["blocksize"]]
    return vals.astype(np.float32).reshape(state["shape"])
