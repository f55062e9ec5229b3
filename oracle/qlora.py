"""QLoRA linear forward / backward, restated on the CPU with torch (fp32 or bf16-emulated).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.

Restates (third-party):

* ``bitsandbytes.autograd._functions.MatMul4Bit``:
    fwd  ``Y0 = X @ dequant(W4).T``  (fp32 accumulate, output in the compute dtype)
    bwd  ``dX0 = dY @ dequant(W4)``  (the weight is decoded again; no dW)
* ``peft.tuners.lora.bnb.Linear4bit.forward`` with one active adapter:
    ``Y = Y0 + lora_B(lora_A(dropout(X))) * (lora_alpha / r)``
  and its autograd backward:
    ``dV = s*dY``, ``dB = dV.T @ U``, ``dU = dV @ B``, ``dA = dU.T @ Xd``,
    ``dX = dX0 + (dU @ A) * mask/(1-p)``.

``mode='bf16'`` rounds to bf16 at every point where the reference (bf16 autocast,
bf16 LoRA weights after cullavo/load_cullavo.py:124-126) materialises a bf16 tensor;
``mode='fp32'`` keeps everything fp32 except the decoded weight, which is the bf16
value bitsandbytes produces for ``compute_dtype=bf16`` (or the unrounded fp32 product
when ``w_bf16=False``, i.e. BASELINE config C1's fp32 compute dtype).
"""
from __future__ import annotations

import numpy as np
import torch

from . import nf4


def _decode(state: dict, w_bf16: bool = True) -> torch.Tensor:
    if w_bf16:
        return torch.from_numpy(nf4.dequantize_nf4(state, as_bits=False).copy())
    return torch.from_numpy(nf4.dequantize_nf4_f32(state).copy())


def _rnd(mode: str):
    if mode == "bf16":
        return lambda t: t.to(torch.bfloat16).to(torch.float32)
    if mode == "fp32":
        return lambda t: t
    raise ValueError(mode)


def qlora_linear_fwd(x, state, A, B, s, mask=None, p=0.0, mode="fp32", w_bf16=True):
    """Returns (y, u, xd).  x [M,K], A [r,K], B [N,r]; mask [M,K] of {0,1} or None."""
    rnd = _rnd(mode)
    x = x.float()
    A = A.float()
    B = B.float()
    W = _decode(state, w_bf16)
    y0 = rnd(x @ W.t())
    if mask is not None:
        xd = rnd(x * mask.float() * (1.0 / (1.0 - p)))
    else:
        xd = x
    u = rnd(xd @ A.t())
    v = rnd(u @ B.t())
    y = rnd(y0 + rnd(v * s))
    return y, u, xd


def qlora_linear_fwd_bwd(x, state, A, B, s, dy, mask=None, p=0.0, mode="fp32", w_bf16=True):
    """One forward + backward.  Returns dict(y, u, dx, dA, dB, du).

    The weight is decoded in both passes, as MatMul4Bit does.
    """
    rnd = _rnd(mode)
    y, u, xd = qlora_linear_fwd(x, state, A, B, s, mask, p, mode, w_bf16)
    dy = dy.float()
    A = A.float()
    B = B.float()
    W = _decode(state, w_bf16)  # second decode (backward)
    dv = rnd(dy * s)
    dB = rnd(dv.t() @ u)
    du = rnd(dv @ B)
    dA = rnd(du.t() @ xd)
    dxl = rnd(du @ A)
    if mask is not None:
        dxl = rnd(dxl * mask.float() * (1.0 / (1.0 - p)))
    dx0 = rnd(dy @ W)
    dx = rnd(dx0 + dxl)
    return {"y": y, "u": u, "dx": dx, "dA": dA, "dB": dB, "du": du}


def qlora_flops(M: int, N: int, K: int, r: int) -> int:
    """Algorithmic fwd+bwd FLOPs of one QLoRA linear; no recompute."""
    return 4 * M * N * K + 6 * M * r * (N + K)


def make_case(M, N, K, r, seed=0, double_quant=True, lora_b_zero=False, p=0.0, dtype=torch.bfloat16):
    """Seeded synthetic case per SURVEY.md section 8d: W~N(0,.02^2) NF4-quantised, X~N(0,1),
    dY~N(0,1)/sqrt(N), A~U(+-1/sqrt(K)), B~N(0,.02^2), mask~Bernoulli(1-p)."""
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(N, K, generator=g) * 0.02
    state = nf4.quantize_nf4(W.numpy(), 64, double_quant)
    x = torch.randn(M, K, generator=g).to(dtype)
    dy = (torch.randn(M, N, generator=g) / (N**0.5)).to(dtype)
    bound = 1.0 / (K**0.5)
    A = ((torch.rand(r, K, generator=g) * 2 - 1) * bound).to(dtype)
    B = (torch.zeros(N, r) if lora_b_zero else torch.randn(N, r, generator=g) * 0.02).to(dtype)
    mask = None
    if p > 0:
        mask = (torch.rand(M, K, generator=torch.Generator().manual_seed(seed + 1)) >= p).to(torch.uint8)
    return {"state": state, "x": x, "dy": dy, "A": A, "B": B, "mask": mask, "p": p}


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Per-tensor max|a-b| / max|b| (primary definition)."""
    a = a.float()
    b = b.float()
    denom = b.abs().max().item()
    if denom == 0:
        return float((a - b).abs().max().item())
    return float((a - b).abs().max().item() / denom)


def state_to_numpy_checksum(state: dict) -> int:
    """Order-sensitive 64-bit checksum of the decoded bf16 bit patterns (for golden files)."""
    bits = nf4.dequantize_nf4(state).reshape(-1).astype(np.uint64)
    idx = np.arange(1, bits.size + 1, dtype=np.uint64)
    return int((bits * idx).sum() & np.uint64(0xFFFFFFFFFFFFFFFF))
