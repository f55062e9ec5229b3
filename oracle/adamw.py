"""AdamW step + global-norm gradient clipping, restated on the CPU (fp32 math, chosen storage dtype).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates what the reference runs on its trainable parameters:
``accel.clip_grad_norm_(model.parameters(), GRAD_MAX)`` (/root/reference/pipeline/CuLLaVOPipeline.py:90-91 ->
``torch.nn.utils.clip_grad_norm_``: ``coef = min(1, max_norm / (||g||_2 + 1e-6))``) followed by
``torch.optim.AdamW.step`` (/root/reference/trainer/cullavo_trainer.py:13, trainer/default_trainer.py:86-90):

    p *= 1 - lr*wd ;  m = b1*m + (1-b1)*g ;  v = b2*v + (1-b2)*g*g
    p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)

This file is pinned against ``torch.optim.AdamW`` itself (tests/test_optim.py) -- torch IS importable here, so unlike
the NF4 part this piece of the oracle is pinned by the real implementation.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch


def clip_coef(grads: Sequence[torch.Tensor], max_norm: float) -> float:
    total = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads))
    return min(1.0, max_norm / (total + 1e-6))


def adamw_step(p, g, m, v, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
               clip: float = 1.0, param_dtype=torch.bfloat16, state_dtype=torch.bfloat16):
    """One step; inputs hold values representable in their storage dtypes.  Returns (p, m, v) rounded to storage."""
    b1, b2 = betas
    p32, g32, m32, v32 = (t.float() for t in (p, g, m, v))
    g32 = g32 * torch.tensor(clip, dtype=torch.float32)
    p32 = p32 * torch.tensor(1.0 - lr * weight_decay, dtype=torch.float32)
    m32 = torch.tensor(b1, dtype=torch.float32) * m32 + torch.tensor(1.0 - b1, dtype=torch.float32) * g32
    v32 = torch.tensor(b2, dtype=torch.float32) * v32 + torch.tensor(1.0 - b2, dtype=torch.float32) * g32 * g32
    inv_bc1 = torch.tensor(1.0 / (1.0 - b1 ** step), dtype=torch.float32)
    inv_sqrt_bc2 = torch.tensor(1.0 / math.sqrt(1.0 - b2 ** step), dtype=torch.float32)
    denom = v32.sqrt() * inv_sqrt_bc2 + torch.tensor(eps, dtype=torch.float32)
    p32 = p32 - (torch.tensor(lr, dtype=torch.float32) * inv_bc1) * (m32 / denom)
    return p32.to(param_dtype), m32.to(state_dtype), v32.to(state_dtype)
