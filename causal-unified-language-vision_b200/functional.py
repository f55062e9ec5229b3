"""Tensor-level ops of the QLoRA linear hot path: thin, validating wrappers over the C ABI.

Names and argument meaning follow ``bitsandbytes.functional`` (``quantize_4bit``,
``dequantize_4bit``, ``QuantState``) where the reference's dependency has an equivalent
(SURVEY.md section 2.2 rows T2-T4); the fused QLoRA ops have no upstream equivalent and are
named after what they compute.  Everything here runs on the GPU through libb2q.so; a CPU
tensor is an error, not a fallback.
"""
from __future__ import annotations

import ctypes as ct
import functools
import os
import threading
from typing import Optional

import torch

from . import _lib

NF4_CODE = [
    -1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
    -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
    0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
    0.7229568362236023, 1.0,
]


def create_dynamic_map(signed: bool = True, max_exponent_bits: int = 7, total_bits: int = 8) -> torch.Tensor:
    """bitsandbytes' 8-bit dynamic code (used for the nested absmax); 256 sorted fp32 values."""
    data = []
    non_sign_bits = total_bits - 1
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    i = 0
    for i in range(max_exponent_bits):
        fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed
                             else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        boundaries = torch.linspace(0.1, 1, fraction_items)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    if additional_items > 0:
        boundaries = torch.linspace(0.1, 1, additional_items + 1)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    data.append(0)
    data.append(1.0)
    assert len(data) == 2**total_bits
    data.sort()
    return torch.tensor(data, dtype=torch.float32)


def _p(t: Optional[torch.Tensor]):
    return ct.c_void_p(0) if t is None else ct.c_void_p(t.data_ptr())


def _stream():
    return ct.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    """Every tensor on the GPU, and on the calling thread's CURRENT device: the C ABI launches on the current device's
    stream (bitsandbytes asserts the same with ``is_on_gpu`` and switches with ``pre_call``; the autograd Functions in
    autograd.py do the switch, direct callers must)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("b2q ops run on the GPU only (sm_100a); got a CPU tensor and there is no CPU fallback")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"tensor on {t.device} but the current CUDA device is cuda:{cur}: wrap the call in "
                               "`with torch.cuda.device(tensor.device):`")


def _on_device_of_first_arg(fn):
    """bitsandbytes' ``pre_call(A.device)`` / ``post_call`` for the functions that mirror its public API: run on the device
    of the first tensor even when that is not the calling thread's current device."""
    @functools.wraps(fn)
    def wrapper(A, *args, **kwargs):
        if isinstance(A, torch.Tensor) and A.is_cuda and A.device.index != torch.cuda.current_device():
            with torch.cuda.device(A.device):
                return fn(A, *args, **kwargs)
        return fn(A, *args, **kwargs)
    return wrapper


def _need(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if t.data_ptr() % 16 != 0:
        raise ValueError(f"{name} must be 16-byte aligned")


class QuantState:
    """Quantisation state of one NF4 weight; field names follow ``bitsandbytes.functional.QuantState``.

    absmax   fp32 [nblocks] (plain)  or  uint8 [nblocks] (nested, then ``state2`` and ``offset`` are set)
    code     fp32 [16]   shape  dtype  blocksize  quant_type
    state2   QuantState(absmax=fp32 [ceil(nblocks/256)], code=fp32 [256], blocksize=256)
    """

    valid_qs_keys = ["absmax", "quant_map", "nested_absmax", "nested_quant_map", "quant_state"]

    def __init__(self, absmax, shape=None, code=None, blocksize=64, quant_type="nf4", dtype=torch.bfloat16,
                 offset=None, state2=None):
        self.absmax = absmax
        self.shape = None if shape is None else torch.Size(shape)
        self.code = code
        self.blocksize = blocksize
        self.quant_type = quant_type
        self.dtype = dtype
        self.offset = offset
        self.state2 = state2
        self.nested = state2 is not None

    def to(self, device):
        self.absmax = self.absmax.to(device)
        self.code = self.code.to(device)
        if self.nested:
            self.offset = self.offset.to(device)
            self.state2.absmax = self.state2.absmax.to(device)
            self.state2.code = self.state2.code.to(device)
        return self

    def as_dict(self, packed: bool = False) -> dict:
        """bitsandbytes' serialisation: tensors + a packed ``quant_state.bitsandbytes__nf4`` blob."""
        qs = {
            "quant_type": self.quant_type,
            "absmax": self.absmax,
            "blocksize": self.blocksize,
            "quant_map": self.code,
            "dtype": str(self.dtype).replace("torch.", ""),
            "shape": tuple(self.shape),
        }
        if self.nested:
            qs.update({
                "nested_absmax": self.state2.absmax,
                "nested_blocksize": self.state2.blocksize,
                "nested_quant_map": self.state2.code.clone(),
                "nested_dtype": "float32",
                "nested_offset": float(self.offset.item()),
            })
        if not packed:
            return qs
        import json

        tensors = {k: v for k, v in qs.items() if isinstance(v, torch.Tensor)}
        non_tensor = {k: v for k, v in qs.items() if not isinstance(v, torch.Tensor)}
        blob = torch.tensor(list(json.dumps(non_tensor).encode("utf-8")), dtype=torch.uint8)
        tensors["quant_state.bitsandbytes__" + self.quant_type] = blob
        return tensors

    @classmethod
    def from_dict(cls, qs_dict: dict, device) -> "QuantState":
        import json

        qs_key = [k for k in qs_dict if "quant_state" in k and isinstance(qs_dict[k], torch.Tensor)]
        d = {k.split(".")[-1]: v for k, v in qs_dict.items()}
        if qs_key:
            blob = qs_dict[qs_key[0]]
            d.update(json.loads(bytes(blob.cpu().tolist()).decode("utf-8")))
            d.pop(qs_key[0].split(".")[-1], None)
        state2 = None
        offset = None
        if "nested_absmax" in d:
            offset = torch.tensor(float(d["nested_offset"]), dtype=torch.float32, device=device)
            state2 = cls(absmax=d["nested_absmax"].to(device), code=d["nested_quant_map"].to(device),
                         blocksize=int(d["nested_blocksize"]), quant_type="dynamic8", dtype=torch.float32)
        return cls(absmax=d["absmax"].to(device), shape=torch.Size(d["shape"]), code=d["quant_map"].to(device),
                   blocksize=int(d["blocksize"]), quant_type=d["quant_type"], dtype=getattr(torch, d["dtype"]),
                   offset=offset, state2=state2)

    # -- C ABI view -----------------------------------------------------------------------
    def c_weight(self, packed: torch.Tensor) -> "_lib.NF4Weight":
        w = _lib.NF4Weight()
        w.packed = packed.data_ptr()
        w.code16 = self.code.data_ptr()
        if self.nested:
            w.absmax = None
            w.absmax_q = self.absmax.data_ptr()
            w.absmax2 = self.state2.absmax.data_ptr()
            w.code256 = self.state2.code.data_ptr()
            if not hasattr(self, "_offset_host"):
                self._offset_host = float(self.offset.item())  # one sync, at first use
            w.offset = self._offset_host
        else:
            w.absmax = self.absmax.data_ptr()
            w.absmax_q = None
            w.absmax2 = None
            w.code256 = None
            w.offset = 0.0
        return w


@_on_device_of_first_arg
def quantize_4bit(A: torch.Tensor, blocksize: int = 64, compress_statistics: bool = False, quant_type: str = "nf4",
                  quant_storage=torch.uint8):
    """NF4 blockwise quantisation on the GPU.  Returns (packed uint8 [(n+1)//2, 1], QuantState).

    Same contract as ``bitsandbytes.functional.quantize_4bit`` for ``quant_type='nf4'``,
    ``blocksize=64`` (the configuration cullavo/load_cullavo.py:73-82 selects).
    """
    if quant_type != "nf4" or blocksize != 64 or quant_storage != torch.uint8:
        raise NotImplementedError("only quant_type='nf4', blocksize=64, quant_storage=uint8 (the reference's config)")
    _need_cuda(A)
    if A.dtype not in (torch.float32, torch.bfloat16):
        A = A.float()
    A = A.contiguous()
    n = A.numel()
    lib = _lib.load()
    dev = A.device
    packed = torch.empty(((n + 1) // 2, 1), dtype=torch.uint8, device=dev)
    nblocks = (n + 63) // 64
    absmax = torch.empty(nblocks, dtype=torch.float32, device=dev)
    code = torch.tensor(NF4_CODE, dtype=torch.float32, device=dev)
    _lib.check(lib.b2q_nf4_quantize(_p(A), int(A.dtype == torch.bfloat16), n, _p(packed), _p(absmax), _stream()),
               "b2q_nf4_quantize")
    if not compress_statistics:
        return packed, QuantState(absmax, A.shape, code, 64, "nf4", A.dtype if A.dtype != torch.float32 else torch.float32)
    code256 = create_dynamic_map().to(dev)
    absmax_q = torch.empty(nblocks, dtype=torch.uint8, device=dev)
    absmax2 = torch.empty((nblocks + 255) // 256, dtype=torch.float32, device=dev)
    offset = torch.empty((), dtype=torch.float32, device=dev)
    _lib.check(lib.b2q_absmax_double_quant(_p(absmax), nblocks, _p(code256), _p(absmax_q), _p(absmax2), _p(offset),
                                           _stream()), "b2q_absmax_double_quant")
    state2 = QuantState(absmax2, code=code256, blocksize=256, quant_type="dynamic8", dtype=torch.float32)
    return packed, QuantState(absmax_q, A.shape, code, 64, "nf4", A.dtype, offset=offset, state2=state2)


@_on_device_of_first_arg
def dequantize_4bit(A: torch.Tensor, quant_state: QuantState, algo: int = 1) -> torch.Tensor:
    """Decode packed NF4 to a bf16 tensor of ``quant_state.shape`` (bit-exact with the reference decode).

    Debug / checkpoint-export utility: the training path never materialises the bf16 weight.
    """
    _need_cuda(A)
    lib = _lib.load()
    n = 1
    for s in quant_state.shape:
        n *= int(s)
    out = torch.empty(tuple(quant_state.shape), dtype=torch.bfloat16, device=A.device)
    w = quant_state.c_weight(A)
    _lib.check(lib.b2q_nf4_decode(w.packed, w.absmax, w.absmax_q, w.absmax2, w.code256, w.offset, w.code16, _p(out), n,
                                  quant_state.blocksize, algo, _stream()), "b2q_nf4_decode")
    return out


GEMV_MAX_ROWS = 8


@_on_device_of_first_arg
def gemv_4bit(x: torch.Tensor, packed: torch.Tensor, qs: QuantState) -> torch.Tensor:
    """y = x @ dequant(W)^T for up to 8 token rows (inference / `generate`): HBM-bound SIMT kernel, no tensor cores.
    Stand-in for ``bitsandbytes.functional.gemv_4bit``."""
    _need_cuda(x, packed)
    _need(x, torch.bfloat16, "x")
    M, K = x.shape
    N = int(qs.shape[0])
    if int(qs.shape[1]) != K:
        raise ValueError(f"x has {K} features, weight expects {int(qs.shape[1])}")
    if M > GEMV_MAX_ROWS:
        raise ValueError(f"gemv_4bit handles at most {GEMV_MAX_ROWS} rows, got {M}")
    y = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    w = qs.c_weight(packed)
    _lib.check(_lib.load().b2q_gemv_4bit(_p(x), ct.byref(w), _p(y), M, N, K, _stream()), "b2q_gemv_4bit")
    return y


# ------------------------------------------------------------------------------ dropout ----
def dropout_mask(shape, seed: int, p: float, device) -> torch.Tensor:
    mask = torch.empty(shape, dtype=torch.uint8, device=device)
    _need_cuda(mask)
    _lib.check(_lib.load().b2q_dropout_mask(_p(mask), mask.numel(), seed, p, _stream()), "b2q_dropout_mask")
    return mask


def dropout_apply(x: torch.Tensor, seed: int, p: float) -> torch.Tensor:
    _need_cuda(x)
    _need(x, torch.bfloat16, "x")
    out = torch.empty_like(x)
    _lib.check(_lib.load().b2q_dropout_apply(_p(x), _p(out), x.numel(), seed, p, _stream()), "b2q_dropout_apply")
    return out


def dropout_bwd_add_(dx: torch.Tensor, dxl: torch.Tensor, seed: int, p: float) -> torch.Tensor:
    _need_cuda(dx, dxl)
    _need(dx, torch.bfloat16, "dx")
    _need(dxl, torch.bfloat16, "dxl")
    _lib.check(_lib.load().b2q_dropout_bwd_add(_p(dx), _p(dxl), dx.numel(), seed, p, _stream()), "b2q_dropout_bwd_add")
    return dx


# --------------------------------------------------------------------------- QLoRA GEMMs ----
def lora_down(x: torch.Tensor, lora_A: torch.Tensor, scale: float, seed: int = 0, p: float = 0.0):
    """u = drop(x) @ A^T, us = scale * u   (x [M,K], A [r,K]) -> (u [M,r], us [M,r]) bf16.
    The dropout mask keep(seed, i)/(1-p) is applied to the x tile in shared memory."""
    _need_cuda(x, lora_A)
    _need(x, torch.bfloat16, "x")
    _need(lora_A, torch.bfloat16, "lora_A")
    M, K = x.shape
    r = lora_A.shape[0]
    u = torch.empty((M, r), dtype=torch.bfloat16, device=x.device)
    us = torch.empty((M, r), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().b2q_lora_down(_p(x), _p(lora_A), scale, seed, p, _p(u), _p(us), M, K, r, _stream()),
               "b2q_lora_down")
    return u, us


def qlora_fwd(x: torch.Tensor, packed: torch.Tensor, qs: QuantState, us: Optional[torch.Tensor],
              lora_B: Optional[torch.Tensor]) -> torch.Tensor:
    """y = x @ dequant(W)^T (+ us @ B^T) in one tcgen05 kernel.  x [M,K] bf16 -> y [M,N] bf16."""
    _need_cuda(x, packed, us, lora_B)
    _need(x, torch.bfloat16, "x")
    M, K = x.shape
    N = int(qs.shape[0])
    if int(qs.shape[1]) != K:
        raise ValueError(f"x has {K} features, weight expects {int(qs.shape[1])}")
    r = 0
    if us is not None:
        _need(us, torch.bfloat16, "us")
        _need(lora_B, torch.bfloat16, "lora_B")
        r = lora_B.shape[1]
    y = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    w = qs.c_weight(packed)
    _lib.check(_lib.load().b2q_qlora_fwd(_p(x), ct.byref(w), _p(us), _p(lora_B), _p(y), M, N, K, r, _stream()),
               "b2q_qlora_fwd")
    return y


def lora_bwd_du(dy: torch.Tensor, lora_B: torch.Tensor, scale: float, p: float = 0.0) -> torch.Tensor:
    """du = scale / (1 - p) * dy @ B   (dy [M,N], B [N,r]) -> [M,r] bf16: the gradient of the LoRA hidden activation with
    the keep-scale of the LoRA dropout folded in (``qlora_bwd_dx`` and ``lora_grads`` both expect it that way)."""
    _need_cuda(dy, lora_B)
    _need(dy, torch.bfloat16, "dy")
    _need(lora_B, torch.bfloat16, "lora_B")
    M, N = dy.shape
    r = lora_B.shape[1]
    du = torch.empty((M, r), dtype=torch.bfloat16, device=dy.device)
    _lib.check(_lib.load().b2q_lora_bwd_du(_p(dy), _p(lora_B), scale, p, _p(du), M, N, r, _stream()), "b2q_lora_bwd_du")
    return du


def qlora_bwd_dx(dy: torch.Tensor, packed: torch.Tensor, qs: QuantState, du: Optional[torch.Tensor],
                 lora_A: Optional[torch.Tensor], seed: int = 0, p: float = 0.0) -> torch.Tensor:
    """dx = dy @ dequant(W) (+ keep * (du @ A)) with ``du = lora_bwd_du(dy, B, scale, p)``.  dy [M,N] bf16 -> dx [M,K] bf16."""
    _need_cuda(dy, packed, du, lora_A)
    _need(dy, torch.bfloat16, "dy")
    M, N = dy.shape
    K = int(qs.shape[1])
    if int(qs.shape[0]) != N:
        raise ValueError(f"dy has {N} features, weight has {int(qs.shape[0])} rows")
    r = 0
    if du is not None:
        _need(du, torch.bfloat16, "du")
        _need(lora_A, torch.bfloat16, "lora_A")
        r = lora_A.shape[0]
    dx = torch.empty((M, K), dtype=torch.bfloat16, device=dy.device)
    w = qs.c_weight(packed)
    _lib.check(_lib.load().b2q_qlora_bwd_dx(_p(dy), ct.byref(w), _p(du), _p(lora_A), seed, p, _p(dx), M, N, K, r,
                                            _stream()), "b2q_qlora_bwd_dx")
    return dx


# Order of the three backward calls of one module.  lora_bwd_du streams dy through the L2; the dB GEMM inside lora_grads reads
# dy again, so running lora_grads BEFORE qlora_bwd_dx lets it hit the part of dy that is still resident (a 134 MB dy against
# 126 MB of L2) instead of re-reading it from HBM after the decode GEMM has flushed it.  B2Q_GRADS_BEFORE_DX=0: dx first.
GRADS_BEFORE_DX = os.environ.get("B2Q_GRADS_BEFORE_DX", "1") == "1"

_ws_cache: dict = {}
_ws_lock = threading.Lock()   # backward runs on autograd engine threads, one per device


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Split-M partials of ``lora_grads``: one buffer per (device, stream), grown on demand.  Work on one stream is
    ordered, so consecutive modules can share it; the caching allocator keeps a replaced buffer alive until the
    kernels already enqueued on that stream are done with it."""
    key = (device, torch.cuda.current_stream().cuda_stream)
    with _ws_lock:
        ws = _ws_cache.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            _ws_cache[key] = ws
        return ws


def lora_grads(dy, x, u, du, scale: float, dA: torch.Tensor, dB: torch.Tensor, accumulate: bool = False,
               seed: int = 0, p: float = 0.0):
    """dA[r,K] (+)= du^T @ (keep * x) ; dB[N,r] (+)= scale * dy^T @ u, with ``du = lora_bwd_du(dy, B, scale, p)`` (keep-scale
    included).  dA / dB are written in place (they may be views into a flat gradient bucket)."""
    _need_cuda(dy, x, u, du, dA, dB)
    for t, nm in ((dy, "dy"), (x, "x"), (u, "u"), (du, "du"), (dA, "dA"), (dB, "dB")):
        _need(t, torch.bfloat16, nm)
    M, N = dy.shape
    K = x.shape[1]
    r = u.shape[1]
    lib = _lib.load()
    nbytes = int(lib.b2q_lora_grads_workspace_bytes(M, N, K, r))
    ws = _workspace(nbytes, dy.device)
    _lib.check(lib.b2q_lora_grads(_p(dy), _p(x), _p(u), _p(du), scale, seed, p, _p(dA), _p(dB), int(accumulate), _p(ws),
                                  ws.numel(), M, N, K, r, _stream()), "b2q_lora_grads")
    return dA, dB


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, b_is_kn: bool, alpha: float = 1.0) -> torch.Tensor:
    """alpha * a @ b (b given [K,N]) or alpha * a @ b^T (b given [N,K]) on the tcgen05 pipeline."""
    _need_cuda(a, b)
    _need(a, torch.bfloat16, "a")
    _need(b, torch.bfloat16, "b")
    M, K = a.shape
    N = b.shape[1] if b_is_kn else b.shape[0]
    d = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _lib.check(_lib.load().b2q_gemm_bf16(_p(a), _p(b), int(b_is_kn), alpha, _p(d), M, N, K, _stream()), "b2q_gemm_bf16")
    return d


def set_variant(fwd: int = -1, dx: int = -1) -> None:
    _lib.load().b2q_set_variant(fwd, dx)


def launch_count() -> int:
    return int(_lib.load().b2q_launch_count())
