"""Fused AdamW + global-norm clip over the flat LoRA buckets.

What it replaces in the reference, for the trainable LoRA parameters:
  ``torch.optim.AdamW(self.model.parameters(), lr, weight_decay)``      /root/reference/trainer/cullavo_trainer.py:13
  ``accel.clip_grad_norm_(trainer.model.parameters(), GRAD_MAX)``       /root/reference/pipeline/CuLLaVOPipeline.py:90-91
  ``optimizer.step(); optimizer.zero_grad(); lr_scheduler.step()``      /root/reference/trainer/default_trainer.py:86-90

``FusedLoraAdamW`` is a ``torch.optim.Optimizer`` (so ``CosineAnnealingLR`` at cullavo_trainer.py:14 drives its
``param_groups[0]['lr']`` unchanged).  It re-homes every LoRA ``A`` / ``B`` into flat bf16 parameter buckets laid out
exactly like ``GradSync``'s gradient buckets (``param.data`` becomes a view), keeps both moments flat, and runs ONE
kernel per bucket (14 bytes of HBM traffic per element) after the gradient all-reduce has finished; the clip
coefficient is computed on the device from per-block partial sums, so a step needs no host synchronisation.

The reference's optimizer and clip cover ALL trainable parameters -- besides the LoRA weights that is
``multi_modal_projector`` (/root/reference/cullavo/load_cullavo.py:128-130).  Parameters registered with
``GradSync(extra_params=...)`` are therefore part of this optimizer too: their squared gradient norm enters the global
norm, the same clip coefficient scales their gradients, and a stock ``torch.optim.AdamW`` with the same
hyper-parameters (sharing ``param_groups[0]['lr']``, so one scheduler drives both) steps them.
"""
from __future__ import annotations

import ctypes as ct
from typing import Optional

import torch

from . import _lib
from .functional import _need_cuda, _p, _stream
from .parallel import GradSync


class FusedLoraAdamW(torch.optim.Optimizer):
    def __init__(self, sync: GradSync, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, state_dtype: torch.dtype = torch.bfloat16):
        if sync.grad_dtype != torch.bfloat16:
            raise ValueError("FusedLoraAdamW works on bf16 gradient buckets")
        if state_dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("state_dtype must be bfloat16 (the reference's: states follow the bf16 params) or float32")
        self.sync = sync
        params = [prm for (_, prm, _) in sync.slots]
        for prm in params:
            if prm.dtype != torch.bfloat16:
                raise TypeError("LoRA parameters must be bf16 (the reference's fp32->bf16 sweep, load_cullavo.py:124-126, "
                                "runs before the optimizer is built)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        # flat parameter / moment buckets mirroring the gradient buckets
        self.pflat, self.mflat, self.vflat = [], [], []
        for b in sync.buckets:
            pf = torch.empty_like(b.flat)
            for i, off in zip(b.members, b.offsets):
                prm = sync.slots[i][1]
                view = pf[off:off + prm.numel()].view_as(prm)
                view.copy_(prm.data)
                prm.data = view
            self.pflat.append(pf)
            self.mflat.append(torch.zeros(b.flat.numel(), dtype=state_dtype, device=b.flat.device))
            self.vflat.append(torch.zeros(b.flat.numel(), dtype=state_dtype, device=b.flat.device))
        self.state_dtype = state_dtype
        self.extra_params = list(sync.extra_params)
        self.extra_optimizer = (torch.optim.AdamW(self.extra_params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
                                if self.extra_params else None)
        self._extra_coef: Optional[torch.Tensor] = None
        self._step = 0
        self._max_norm = 0.0
        self._partials: Optional[torch.Tensor] = None
        self._partials_valid = False

    # ---- gradient clipping ---------------------------------------------------------------
    def _compute_partials(self) -> torch.Tensor:
        lib = _lib.load()
        counts = [int(lib.b2q_sqnorm_blocks(b.flat.numel())) for b in self.sync.buckets]
        total = sum(counts) + 1   # last slot: squared gradient norm of the non-LoRA trainable parameters (0 if none)
        dev = self.sync.device
        if self._partials is None or self._partials.numel() != total:
            self._partials = torch.zeros(total, dtype=torch.float32, device=dev)
        grads = [p.grad for p in self.extra_params if p.grad is not None]
        if grads:
            self._partials[total - 1] = torch.stack([g.float().pow(2).sum() for g in grads]).sum()
        else:
            self._partials[total - 1] = 0.0
        off = 0
        for b, c in zip(self.sync.buckets, counts):
            _need_cuda(b.flat)
            _lib.check(lib.b2q_sqnorm_partials(_p(b.flat), b.flat.numel(),
                                               ct.c_void_p(self._partials.data_ptr() + 4 * off), _stream()),
                       "b2q_sqnorm_partials")
            off += c
        self._partials_valid = True
        return self._partials

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """Arms the global-norm clip for the next ``step()`` and returns the total gradient norm (device scalar, no
        host sync).  Same formula as ``torch.nn.utils.clip_grad_norm_``; the scaling itself is fused into the step."""
        self.sync.finish()  # the all-reduced gradients are what gets clipped
        self._max_norm = float(max_norm)
        norm = self._compute_partials().sum().sqrt()
        if self.extra_params:   # the fused kernel scales the LoRA gradients; the others are scaled here, same coefficient
            self._extra_coef = torch.clamp(self._max_norm / (norm + 1e-6), max=1.0)
        return norm

    # ---- torch.optim.Optimizer surface ---------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.sync.finish()
        lib = _lib.load()
        g = self.param_groups[0]
        self._step += 1
        partials = None
        if self._max_norm > 0.0:
            partials = self._partials if self._partials_valid else self._compute_partials()
        for b, pf, mf, vf in zip(self.sync.buckets, self.pflat, self.mflat, self.vflat):
            _lib.check(lib.b2q_adamw_step(_p(pf), _p(b.flat), _p(mf), _p(vf), int(self.state_dtype == torch.float32),
                                          pf.numel(), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                                          float(g["eps"]), float(g["weight_decay"]), self._step, _p(partials),
                                          0 if partials is None else partials.numel(), self._max_norm, _stream()),
                       "b2q_adamw_step")
        if self.extra_optimizer is not None:
            if self._max_norm > 0.0 and self._extra_coef is not None:
                grads = [p.grad for p in self.extra_params if p.grad is not None]
                if grads:
                    torch._foreach_mul_(grads, self._extra_coef)
            self.extra_optimizer.param_groups[0]["lr"] = g["lr"]
            self.extra_optimizer.step()
        self._extra_coef = None
        self._partials_valid = False
        self._max_norm = 0.0  # like the reference, clipping is requested per step
        return loss

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Gradients live in the buckets; they are zeroed (never detached) so the views stay valid."""
        self.sync.zero_grad()
        if self.extra_optimizer is not None:
            self.extra_optimizer.zero_grad(set_to_none=True)

    def state_dict(self):
        sd = super().state_dict()
        sd["b2q_flat"] = {"step": self._step, "m": [t.clone() for t in self.mflat], "v": [t.clone() for t in self.vflat]}
        if self.extra_optimizer is not None:
            sd["b2q_extra"] = self.extra_optimizer.state_dict()
        return sd

    def load_state_dict(self, state_dict):
        flat = state_dict.get("b2q_flat")
        if self.extra_optimizer is not None and "b2q_extra" in state_dict:
            self.extra_optimizer.load_state_dict(state_dict["b2q_extra"])
        super().load_state_dict({k: v for k, v in state_dict.items() if k not in ("b2q_flat", "b2q_extra")})
        if flat is not None:
            self._step = int(flat["step"])
            for dst, src in zip(self.mflat, flat["m"]):
                dst.copy_(src)
            for dst, src in zip(self.vflat, flat["v"]):
                dst.copy_(src)
