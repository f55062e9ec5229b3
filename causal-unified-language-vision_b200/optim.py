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
        self._step = 0
        self._max_norm = 0.0
        self._partials: Optional[torch.Tensor] = None
        self._partials_valid = False

    # ---- gradient clipping ---------------------------------------------------------------
    def _compute_partials(self) -> torch.Tensor:
        lib = _lib.load()
        counts = [int(lib.b2q_sqnorm_blocks(b.flat.numel())) for b in self.sync.buckets]
        total = sum(counts)
        dev = self.sync.device
        if self._partials is None or self._partials.numel() != total:
            self._partials = torch.zeros(total, dtype=torch.float32, device=dev)
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
        return self._compute_partials().sum().sqrt()

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
        self._partials_valid = False
        self._max_norm = 0.0  # like the reference, clipping is requested per step
        return loss

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Gradients live in the buckets; they are zeroed (never detached) so the views stay valid."""
        self.sync.zero_grad()

    def state_dict(self):
        sd = super().state_dict()
        sd["b2q_flat"] = {"step": self._step, "m": [t.clone() for t in self.mflat], "v": [t.clone() for t in self.vflat]}
        return sd

    def load_state_dict(self, state_dict):
        flat = state_dict.get("b2q_flat")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "b2q_flat"})
        if flat is not None:
            self._step = int(flat["step"])
            for dst, src in zip(self.mflat, flat["m"]):
                dst.copy_(src)
            for dst, src in zip(self.vflat, flat["v"]):
                dst.copy_(src)
