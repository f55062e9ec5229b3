"""The "QLoRA linear stack" workload: the 7 projections of each decoder layer, forward for all
then backward for all, on fixed synthetic activations (BASELINE.json configs C2-C5; SURVEY.md
section 8d).  No attention / norm / activation functions in between -- the metric is
"linear-stack" tokens/s.

Two drivers over the same modules and weights:
  * ``step_direct``  calls the C-ABI ops back to back with device-resident inputs (bench ``value``);
  * ``step_modules`` goes through ``LoraLinear4bit.forward`` + autograd, i.e. the reference-facing
    plugin surface (bench ``e2e``, which adds the host<->device copies around it).
Both write LoRA gradients straight into ``GradSync`` buckets and overlap the all-reduce.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import functional as F
from .lora import LoraLinear4bit
from .nn import Linear4bit, Params4bit
from .parallel import GradSync

# (name, out_features N, in_features K) -- BASELINE.json's literal "Mistral-7B-shaped" layer
MISTRAL_LITERAL = [("q_proj", 4096, 4096), ("k_proj", 4096, 4096), ("v_proj", 4096, 4096), ("o_proj", 4096, 4096),
                   ("gate_proj", 14336, 4096), ("up_proj", 14336, 4096), ("down_proj", 4096, 14336)]
MISTRAL_GQA = [("q_proj", 4096, 4096), ("k_proj", 1024, 4096), ("v_proj", 1024, 4096), ("o_proj", 4096, 4096),
               ("gate_proj", 14336, 4096), ("up_proj", 14336, 4096), ("down_proj", 4096, 14336)]
LLAMA_7B = [("q_proj", 4096, 4096), ("k_proj", 4096, 4096), ("v_proj", 4096, 4096), ("o_proj", 4096, 4096),
            ("gate_proj", 11008, 4096), ("up_proj", 11008, 4096), ("down_proj", 4096, 11008)]
SHAPE_SETS = {"mistral_literal": MISTRAL_LITERAL, "mistral_gqa": MISTRAL_GQA, "llama_7b": LLAMA_7B}


def stack_flops_per_token(shapes: Sequence[Tuple[str, int, int]], n_layers: int, r: int) -> int:
    """Algorithmic fwd+bwd FLOPs per token: sum over linears of 4NK + 6r(N+K) (no recompute)."""
    return n_layers * sum(4 * n * k + 6 * r * (n + k) for _, n, k in shapes)


def make_quantized_linear(N: int, K: int, device, gen: torch.Generator, double_quant: bool = True) -> Linear4bit:
    """``Linear4bit`` whose weight ~ N(0, 0.02^2) is drawn and NF4-quantised on the GPU (no host fp32 copy)."""
    lin = Linear4bit(K, N, bias=False, compute_dtype=torch.bfloat16, compress_statistics=double_quant,
                     quant_type="nf4", device="meta")
    w = torch.empty(N, K, device=device, dtype=torch.float32).normal_(0.0, 0.02, generator=gen)
    packed, qs = F.quantize_4bit(w, compress_statistics=double_quant)
    del w
    lin.weight = Params4bit(packed, requires_grad=False, quant_state=qs, compress_statistics=double_quant,
                            quant_type="nf4", module=lin, bnb_quantized=True)
    lin.quant_state = qs
    return lin


class QLoRALinearStack(nn.Module):
    def __init__(self, n_layers: int, shapes: Sequence[Tuple[str, int, int]], M: int, r: int = 64, lora_alpha: int = 16,
                 dropout: float = 0.05, double_quant: bool = True, device="cuda:0", seed: int = 0,
                 adapter: str = "step1", bucket_bytes: int = 64 << 20, process_group=None):
        super().__init__()
        self.M, self.r, self.adapter, self.p = M, r, adapter, float(dropout)
        self.shapes = list(shapes)
        self.n_layers = n_layers
        dev = torch.device(device)
        gen = torch.Generator(device=dev).manual_seed(seed)
        self.layers = nn.ModuleList()
        for _ in range(n_layers):
            layer = nn.ModuleDict()
            for name, N, K in self.shapes:
                lin = make_quantized_linear(N, K, dev, gen, double_quant)
                mod = LoraLinear4bit(lin, adapter, r=r, lora_alpha=lora_alpha, lora_dropout=dropout)
                with torch.no_grad():  # B != 0 so the LoRA branch is exercised
                    mod.lora_B[adapter].weight.normal_(0.0, 0.02)
                layer[name] = mod
            self.layers.append(layer)
        for prm in self.parameters():  # the reference's fp32 -> bf16 sweep (load_cullavo.py:124-126)
            if prm.dtype == torch.float32:
                prm.data = prm.data.to(torch.bfloat16)
        self.mods: List[LoraLinear4bit] = [layer[name] for layer in self.layers for name, _, _ in self.shapes]
        self.sync = GradSync(self.mods, adapter, process_group=process_group, bucket_bytes=bucket_bytes)
        self.sync.broadcast_parameters(0)  # LoRA init draws from the global RNGs: replicas start from rank 0's copy
        # fixed synthetic activations per distinct width (rank-dependent seed for the activations)
        self.inputs = {}
        self.grads_out = {}
        rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        agen = torch.Generator(device=dev).manual_seed(1000 + seed + rank)
        for _, N, K in self.shapes:
            if K not in self.inputs:
                self.inputs[K] = torch.empty(M, K, device=dev, dtype=torch.float32).normal_(generator=agen).bfloat16()
            if N not in self.grads_out:
                g = torch.empty(M, N, device=dev, dtype=torch.float32).normal_(generator=agen)
                self.grads_out[N] = (g / N ** 0.5).bfloat16()
        self.train()
        self._step = 0

    # ------------------------------------------------------------------------------------
    def flops_per_token(self) -> int:
        return stack_flops_per_token(self.shapes, self.n_layers, self.r)

    def _weights(self, mod: LoraLinear4bit):
        base = mod.base_layer
        return (base.weight.data, base.weight.quant_state, mod.lora_A[self.adapter].weight,
                mod.lora_B[self.adapter].weight, mod.scaling[self.adapter])

    def grad_sqnorm(self) -> torch.Tensor:
        """Squared norm of all (reduced) LoRA gradients, a device scalar -- the step's read-back result in bench.py."""
        total = None
        for flat in self.sync.flat_grads():
            v = flat.float().pow(2).sum()
            total = v if total is None else total + v
        return total

    def step_direct(self, recompute: bool = False, inputs: Optional[dict] = None,
                    grads_out: Optional[dict] = None) -> None:
        """Forward for all linears then backward for all, calling the C-ABI ops directly."""
        p = self.p
        inputs = self.inputs if inputs is None else inputs
        grads_out = self.grads_out if grads_out is None else grads_out
        self.sync.begin_step()
        saved = []
        base_seed = 7919 * self._step
        for i, mod in enumerate(self.mods):
            packed, qs, A, B, s = self._weights(mod)
            x = inputs[A.shape[1]]
            seed = base_seed + i
            u, us = F.lora_down(x, A, s, seed, p)
            y = F.qlora_fwd(x, packed, qs, us, B)
            saved.append((u, seed))
            del y, us
        for i in range(len(self.mods) - 1, -1, -1):
            mod = self.mods[i]
            packed, qs, A, B, s = self._weights(mod)
            x = inputs[A.shape[1]]
            dy = grads_out[B.shape[0]]
            u, seed = saved[i]
            if recompute:  # gradient checkpointing re-runs the forward inside backward (load_cullavo.py:91-93)
                u, us = F.lora_down(x, A, s, seed, p)
                y = F.qlora_fwd(x, packed, qs, us, B)
                del y, us
            du = F.lora_bwd_du(dy, B, s, p)
            sink = self.sync.sink_for(mod)
            if F.GRADS_BEFORE_DX:   # dB re-reads dy while lora_bwd_du's pass over it is still in the L2 (functional.py)
                F.lora_grads(dy, x, u, du, s, sink.dA, sink.dB, accumulate=sink.accumulate(), seed=seed, p=p)
                sink.ready()
                dx = F.qlora_bwd_dx(dy, packed, qs, du, A, seed, p)
            else:
                dx = F.qlora_bwd_dx(dy, packed, qs, du, A, seed, p)
                F.lora_grads(dy, x, u, du, s, sink.dA, sink.dB, accumulate=sink.accumulate(), seed=seed, p=p)
                sink.ready()
            del dx, du
        self.sync.finish()
        self._step += 1

    def step_modules(self, inputs: Optional[dict] = None, grads_out: Optional[dict] = None, trace=None,
                     interleaved: bool = False) -> torch.Tensor:
        """Same work through ``LoraLinear4bit.forward`` + autograd.  Returns the squared gradient norm (device scalar)."""
        inputs = self.inputs if inputs is None else inputs
        grads_out = self.grads_out if grads_out is None else grads_out
        self.sync.begin_step()
        if interleaved:
            # memory-lean order (forward + backward per module, last module first): same kernels and launch count as
            # the all-forward / all-backward order, but no 50 GB of live outputs
            for i in range(len(self.mods) - 1, -1, -1):
                mod = self.mods[i]
                x = inputs[mod.in_features].detach().requires_grad_(True)
                y = mod(x)
                torch.autograd.backward(y, grads_out[y.shape[-1]])
                del y, x
                if trace is not None and i % 56 == 0:
                    trace(f"fwd+bwd down to module {i} done")
            self.sync.finish()
            self._step += 1
            return self.grad_sqnorm()
        outs = []
        for mod in self.mods:
            x = inputs[mod.in_features].detach().requires_grad_(True)
            outs.append((mod(x), x))
        if trace is not None:
            trace("forward done")
        for i in range(len(self.mods) - 1, -1, -1):
            y, x = outs[i]
            torch.autograd.backward(y, grads_out[y.shape[-1]])
            outs[i] = None
            if trace is not None and i % 56 == 0:
                trace(f"backward down to module {i} done")
        self.sync.finish()
        if trace is not None:
            trace("grad sync done")
        self._step += 1
        return self.grad_sqnorm()
