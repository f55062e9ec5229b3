"""Autograd glue: the two ``torch.autograd.Function``s that put the C-ABI kernels behind modules.

``MatMul4Bit``   stands in for ``bitsandbytes.autograd._functions.MatMul4Bit`` (frozen NF4 base
                 only): forward ``Y = X W^T``, backward ``dX = dY W``, no dW.
``QLoRALinear``  base + one active LoRA adapter, everything the PEFT wrapper's forward and
                 its autograd backward compute in five
                 launches forward+backward instead of ~25.

Saved for backward: ``x`` (the caller's own tensor), ``u = drop(x) A^T`` [M,r] and the dropout
seed (the mask is regenerated inside the kernels) -- never a decoded weight or a masked copy of x.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch

from . import functional as F


def _device_guard(t: torch.Tensor):
    """bitsandbytes' ``pre_call(A.device)``: the C ABI launches on the current device of the calling thread (backward runs
    on an autograd engine thread).  Non-CUDA tensors pass through so that the functional layer raises its usual error."""
    return torch.cuda.device(t.device) if t.is_cuda else contextlib.nullcontext()


def _flatten(x: torch.Tensor):
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype != torch.bfloat16:
        x2 = x2.to(torch.bfloat16)
    return x2.contiguous(), lead


class MatMul4Bit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, packed, qs):
        with _device_guard(x):
            x2, lead = _flatten(x)
            ctx.qs = qs
            ctx.lead = lead
            ctx.in_dtype = x.dtype
            ctx.save_for_backward(packed)
            y = F.qlora_fwd(x2, packed, qs, None, None)
            return y.reshape(*lead, y.shape[-1])

    @staticmethod
    def backward(ctx, dy):
        (packed,) = ctx.saved_tensors
        dx = None
        if ctx.needs_input_grad[0]:
            with _device_guard(dy):
                dy2, _ = _flatten(dy)
                dx = F.qlora_bwd_dx(dy2, packed, ctx.qs, None, None).reshape(*ctx.lead, -1).to(ctx.in_dtype)
        return dx, None, None


def matmul_4bit(x: torch.Tensor, weight: torch.Tensor, quant_state: F.QuantState,
                bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``bnb.matmul_4bit(x, W.t(), bias, quant_state)`` equivalent (``weight`` is the packed tensor)."""
    packed = weight.data if isinstance(weight, torch.nn.Parameter) else weight
    rows = x.numel() // x.shape[-1] if x.shape[-1] else 0
    if 0 < rows <= F.GEMV_MAX_ROWS and not (x.requires_grad and torch.is_grad_enabled()):
        # bitsandbytes takes its GEMV kernel for single-token inference (no gradient); same rule here
        x2, lead = _flatten(x)
        y = F.gemv_4bit(x2, packed, quant_state).reshape(*lead, -1)
    else:
        y = MatMul4Bit.apply(x, packed, quant_state)
    if bias is not None:
        y = y + bias.to(y.dtype)
    return y


class QLoRALinear(torch.autograd.Function):
    """y = x W^T + s * (drop(x) A^T) B^T  with dX, dA, dB.

    ``grad_sink``: optional ``parallel.GradSink`` -- when given, the LoRA gradients are written by
    the kernels straight into its (bucket) views and ``None`` is returned to autograd for A and B
    (the data-parallel gradient sync owns them; see parallel.py).
    """

    @staticmethod
    def forward(ctx, x, packed, qs, A, B, scale, p, seed, grad_sink):
        with _device_guard(x):
            return QLoRALinear._forward(ctx, x, packed, qs, A, B, scale, p, seed, grad_sink)

    @staticmethod
    def _forward(ctx, x, packed, qs, A, B, scale, p, seed, grad_sink):
        x2, lead = _flatten(x)
        a = A if A.dtype == torch.bfloat16 else A.to(torch.bfloat16)
        b = B if B.dtype == torch.bfloat16 else B.to(torch.bfloat16)
        a = a.contiguous()
        b = b.contiguous()
        u, us = F.lora_down(x2, a, scale, seed, p)   # dropout mask applied in shared memory
        y = F.qlora_fwd(x2, packed, qs, us, b)
        ctx.qs, ctx.lead, ctx.in_dtype = qs, lead, x.dtype
        ctx.scale, ctx.p, ctx.seed, ctx.grad_sink = float(scale), float(p), int(seed), grad_sink
        ctx.param_dtypes = (A.dtype, B.dtype)
        ctx.save_for_backward(x2, packed, a, b, u)
        return y.reshape(*lead, y.shape[-1])

    @staticmethod
    def backward(ctx, dy):
        with _device_guard(dy):
            return QLoRALinear._backward(ctx, dy)

    @staticmethod
    def _backward(ctx, dy):
        x2, packed, a, b, u = ctx.saved_tensors
        dy2, _ = _flatten(dy)
        need_x, need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        du = F.lora_bwd_du(dy2, b, ctx.scale, ctx.p)   # keep-scale of the LoRA dropout folded in

        def input_grad():
            if not need_x:
                return None
            return F.qlora_bwd_dx(dy2, packed, ctx.qs, du, a, ctx.seed, ctx.p).reshape(*ctx.lead, -1).to(ctx.in_dtype)

        def weight_grads():
            if not (need_a or need_b):
                return None, None
            if ctx.grad_sink is not None:
                sink = ctx.grad_sink
                F.lora_grads(dy2, x2, u, du, ctx.scale, sink.dA, sink.dB, accumulate=sink.accumulate(),
                             seed=ctx.seed, p=ctx.p)
                sink.ready()  # may launch the bucket's all-reduce on the comm stream
                return None, None
            gA = torch.empty_like(a)
            gB = torch.empty_like(b)
            F.lora_grads(dy2, x2, u, du, ctx.scale, gA, gB, accumulate=False, seed=ctx.seed, p=ctx.p)
            return (gA.to(ctx.param_dtypes[0]) if need_a else None), (gB.to(ctx.param_dtypes[1]) if need_b else None)

        if F.GRADS_BEFORE_DX:   # dB re-reads dy while lora_bwd_du's pass over it is still in the L2 (functional.py)
            dA, dB = weight_grads()
            dx = input_grad()
        else:
            dx = input_grad()
            dA, dB = weight_grads()
        return dx, None, None, dA, dB, None, None, None, None


def qlora_linear(x, packed, qs, A, B, scale: float, p: float = 0.0, seed: int = 0, grad_sink=None):
    return QLoRALinear.apply(x, packed, qs, A, B, scale, p, seed, grad_sink)
