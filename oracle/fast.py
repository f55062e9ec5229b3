"""Threaded front-end of the C restatement (oracle/nf4_ref.c) for the TIMED CPU baseline of bench.py.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.

The numpy oracle decodes 60 M weights per second, which would make the CPU arm of the bench a measurement of numpy's
fancy indexing rather than of the reference algorithm (bitsandbytes' CPU backend is multi-threaded C++).  This module
runs the same plain-C arithmetic (bit-identical to oracle/nf4.py: tests/test_oracle_c.py) on all host cores by handing
disjoint, block-aligned chunks to `nf4ref_quantize` / `nf4ref_dequantize_bf16` from a thread pool (ctypes releases the
GIL).  It also provides the "one decoder layer" sample that bench.py times.
"""
from __future__ import annotations

import ctypes as ct
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Sequence, Tuple

import numpy as np
import torch

from . import build_c, nf4
from .qlora import qlora_flops

_CHUNK = 64 * 256 * 64  # elements; a multiple of 64 * 256 keeps the nested-absmax indices of a chunk self-contained
_lib = None
_pool = None


def _c():
    global _lib
    if _lib is None:
        lib = ct.CDLL(build_c.build())
        lib.nf4ref_quantize.argtypes = [ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p]
        lib.nf4ref_dequantize_bf16.argtypes = [ct.c_void_p] * 5 + [ct.c_float, ct.c_void_p, ct.c_int64, ct.c_void_p]
        _lib = lib
    return _lib


def _threads(n: int) -> ThreadPoolExecutor:
    global _pool
    if _pool is None or _pool._max_workers != n:
        _pool = ThreadPoolExecutor(max_workers=n)
    return _pool


def _ptr(a: np.ndarray, byte_offset: int = 0):
    return ct.c_void_p(a.ctypes.data + byte_offset)


def quantize_nf4(W: np.ndarray, double_quant: bool, threads: int) -> dict:
    """Same state dict as oracle.nf4.quantize_nf4 (blocksize 64); the per-block part runs in C on `threads` cores."""
    W = np.ascontiguousarray(W, dtype=np.float32)
    n = W.size
    assert n % 64 == 0
    packed = np.empty(n // 2, dtype=np.uint8)
    absmax = np.empty(n // 64, dtype=np.float32)
    lib = _c()

    def job(i0):
        m = min(_CHUNK, n - i0)
        lib.nf4ref_quantize(_ptr(W, 4 * i0), m, _ptr(packed, i0 // 2), _ptr(absmax, 4 * (i0 // 64)))

    list(_threads(threads).map(job, range(0, n, _CHUNK)))
    state = {"packed": packed, "shape": tuple(W.shape), "blocksize": 64, "code": nf4.NF4_CODE.astype(np.float32),
             "nested": bool(double_quant)}
    if not double_quant:
        state["absmax"] = absmax
        return state
    state.update(nf4.double_quantize_absmax(absmax))
    return state


def dequantize_f32(state: dict, threads: int) -> torch.Tensor:
    """fp32 values of the bf16-rounded decode (what MatMul4Bit multiplies with for compute_dtype=bf16)."""
    n = int(np.prod(state["shape"]))
    out = np.empty(n, dtype=np.uint16)
    lib = _c()
    nested = bool(state["nested"])
    code16 = np.ascontiguousarray(state["code"], dtype=np.float32)
    off = float(state["offset"]) if nested else 0.0

    def job(i0):
        m = min(_CHUNK, n - i0)
        b0 = i0 // 64
        if nested:
            lib.nf4ref_dequantize_bf16(_ptr(state["packed"], i0 // 2), None, _ptr(state["absmax_q"], b0),
                                       _ptr(state["absmax2"], 4 * (b0 // 256)), _ptr(state["code256"]), off, _ptr(code16), m,
                                       _ptr(out, 2 * i0))
        else:
            lib.nf4ref_dequantize_bf16(_ptr(state["packed"], i0 // 2), _ptr(state["absmax"], 4 * b0), None, None, None, 0.0,
                                       _ptr(code16), m, _ptr(out, 2 * i0))

    list(_threads(threads).map(job, range(0, n, _CHUNK)))
    w = torch.from_numpy(out.view(np.int16)).view(torch.bfloat16).reshape(state["shape"])
    return w.float()


class LayerSample:
    """One decoder layer of the bench workload on the host: the 7 QLoRA linears (NF4 double-quant base + LoRA), fp32
    forward + backward (dX, dA, dB) on `tokens` rows, the weight decoded in both passes as MatMul4Bit does.  The weights
    are quantised once, outside the timed region (the GPU arm's weights are resident too)."""

    def __init__(self, shapes: Sequence[Tuple[str, int, int]], r: int, tokens: int, threads: int, seed: int = 0):
        self.threads, self.tokens, self.r = threads, tokens, r
        torch.set_num_threads(threads)
        g = torch.Generator().manual_seed(seed)
        self.cases = []
        cache = {}
        for _, N, K in shapes:
            if (N, K) not in cache:   # same-shaped projections share one quantised weight (content does not change the time)
                W = torch.randn(N, K, generator=g) * 0.02
                cache[(N, K)] = quantize_nf4(W.numpy(), True, threads)
                del W
            x = torch.randn(tokens, K, generator=g)
            dy = torch.randn(tokens, N, generator=g) / N ** 0.5
            A = (torch.rand(r, K, generator=g) * 2 - 1) / K ** 0.5
            B = torch.randn(N, r, generator=g) * 0.02
            self.cases.append((cache[(N, K)], x, dy, A, B, N, K))
        self.flops = sum(qlora_flops(tokens, N, K, r) for *_, N, K in self.cases)

    def step(self) -> None:
        s = 16.0 / self.r
        for state, x, dy, A, B, N, K in self.cases:
            W = dequantize_f32(state, self.threads)          # forward decode
            y = x @ W.t() + ((x @ A.t()) @ B.t()) * s
            u = x @ A.t()
            del W
            W = dequantize_f32(state, self.threads)          # backward decodes again (no bf16 copy is kept)
            dv = dy * s
            dB = dv.t() @ u
            du = dv @ B
            dA = du.t() @ x
            dx = dy @ W + du @ A
            del W, y, dB, dA, dx
