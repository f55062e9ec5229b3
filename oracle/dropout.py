"""LoRA-dropout keep mask, restated on the CPU with numpy.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference uses ``nn.Dropout(p=0.05)`` on the LoRA branch input (/root/reference/cullavo/load_cullavo.py:98,107), i.e.
torch's Philox stream; bit-matching that stream is not meaningful, so the product defines its own
counter-based mask (csrc/b2q_internal.h) and this file restates that definition independently of the CUDA code:

    (a, b) = 4 rounds of  { p = a * 0xD2511F53 (64 bit);  a = hi(p) ^ b ^ key;  b = lo(p);  key += 0x9E3779B9 }
             starting from a = i >> 2, b = seed >> 32, key = seed & 0xFFFFFFFF
    field(i) = low 15 bits of the (i & 3)-th 16-bit quarter of (a | b << 32)
    keep(i)  = field(i) >= round(p * 32768)

tests/test_oracle.py pins this file against the host compilation of the product's own ``dropout_keep`` (the very source
the device executes); tests/test_gpu_parity.py checks the mask the GPU exports against it.
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_W = 0x9E3779B9
_MASK32 = np.uint64(0xFFFFFFFF)


def threshold15(p: float) -> int:
    t = int(p * 32768.0 + 0.5)
    return min(t, 32767)


def hash64(seed: int, counters: np.ndarray):
    """counters: uint32 array -> (a, b) uint32 arrays."""
    a = counters.astype(np.uint64)
    b = np.full(a.shape, (seed >> 32) & 0xFFFFFFFF, dtype=np.uint64)
    key = seed & 0xFFFFFFFF
    for _ in range(4):
        prod = a * _M0                      # a < 2^32 and M0 < 2^32: exact in uint64
        a = ((prod >> np.uint64(32)) ^ b ^ np.uint64(key)) & _MASK32
        b = prod & _MASK32
        key = (key + _W) & 0xFFFFFFFF
    return a.astype(np.uint32), b.astype(np.uint32)


def keep_mask(shape, seed: int, p: float) -> np.ndarray:
    """uint8 array of `shape` (row-major element index i): 1 = keep, 0 = drop."""
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.uint64)
    a, b = hash64(int(seed), (idx >> np.uint64(2)).astype(np.uint32))
    word = np.where((idx & np.uint64(2)) != 0, b, a)
    field = np.where((idx & np.uint64(1)) != 0, word >> np.uint32(16), word) & np.uint32(0x7FFF)
    return (field >= np.uint32(threshold15(p))).astype(np.uint8).reshape(shape)
