"""``Params4bit`` / ``Linear4bit``: the module surface the reference instantiates implicitly.

The reference never names these classes; it gets them from HF ``from_pretrained(...,
quantization_config=BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_quant_type='nf4',
bnb_4bit_use_double_quant=True, bnb_4bit_compute_dtype=bf16))`` at
/root/reference/cullavo/load_cullavo.py:73-86, which swaps every ``nn.Linear`` of the LLM for
``bitsandbytes.nn.Linear4bit``.  This file keeps that surface (constructor signature,
attributes, state-dict keys, ``nn.Linear`` subclassing so that ``find_all_linear_names``
at load_cullavo.py:8-20 still finds the projections) on top of the sm_100a kernels.
"""
from __future__ import annotations

import warnings
from typing import Optional

import torch

from . import functional as F
from .autograd import matmul_4bit


class Params4bit(torch.nn.Parameter):
    """Packed NF4 weight (uint8 ``[(N*K+1)//2, 1]``) + its ``QuantState``; quantises on ``.cuda()`` / ``.to('cuda')``.

    Mirrors ``bitsandbytes.nn.Params4bit``.
    """

    def __new__(cls, data: Optional[torch.Tensor] = None, requires_grad: bool = False, quant_state=None,
                blocksize: int = 64, compress_statistics: bool = True, quant_type: str = "fp4",
                quant_storage: torch.dtype = torch.uint8, module=None, bnb_quantized: bool = False):
        if data is None:
            data = torch.empty(0)
        self = torch.Tensor._make_subclass(cls, data, requires_grad)
        self.blocksize = blocksize
        self.compress_statistics = compress_statistics
        self.quant_type = quant_type
        self.quant_state = quant_state
        self.quant_storage = quant_storage
        self.bnb_quantized = bnb_quantized
        self.module = module
        return self

    def __deepcopy__(self, memo):
        import copy

        new = type(self).__new__(type(self), self.data.clone(), self.requires_grad, copy.deepcopy(self.quant_state),
                                 self.blocksize, self.compress_statistics, self.quant_type, self.quant_storage, None,
                                 self.bnb_quantized)
        memo[id(self)] = new
        return new

    @classmethod
    def from_prequantized(cls, data: torch.Tensor, quantized_stats: dict, requires_grad: bool = False,
                          device="cuda", module=None, **kwargs) -> "Params4bit":
        self = torch.Tensor._make_subclass(cls, data.to(device))
        self.requires_grad = requires_grad
        self.quant_state = F.QuantState.from_dict(qs_dict=quantized_stats, device=device)
        self.blocksize = self.quant_state.blocksize
        self.compress_statistics = self.quant_state.nested
        self.quant_type = self.quant_state.quant_type
        self.quant_storage = data.dtype
        self.bnb_quantized = True
        self.module = module
        if module is not None:
            module.quant_state = self.quant_state
        return self

    def _quantize(self, device):
        w = self.data.contiguous().to(device)
        w_4bit, quant_state = F.quantize_4bit(w, blocksize=self.blocksize, compress_statistics=self.compress_statistics,
                                              quant_type=self.quant_type, quant_storage=self.quant_storage)
        self.data = w_4bit
        self.quant_state = quant_state
        if self.module is not None:
            self.module.quant_state = quant_state
        self.bnb_quantized = True
        return self

    def cuda(self, device=None, non_blocking: bool = False):
        return self.to(device="cuda" if device is None else device, non_blocking=non_blocking)

    def to(self, *args, **kwargs):
        device, dtype, non_blocking, _ = torch._C._nn._parse_to(*args, **kwargs)
        if device is not None and device.type == "cuda" and not self.bnb_quantized:
            return self._quantize(device)
        if self.quant_state is not None and device is not None:
            self.quant_state.to(device)
        # the packed bytes never change dtype (the reference's fp32->bf16 sweep at
        # load_cullavo.py:124-126 only touches fp32 parameters, and must not touch these)
        new = Params4bit(super().to(device=device, dtype=None, non_blocking=non_blocking),
                         requires_grad=self.requires_grad, quant_state=self.quant_state, blocksize=self.blocksize,
                         compress_statistics=self.compress_statistics, quant_type=self.quant_type,
                         quant_storage=self.quant_storage, module=self.module, bnb_quantized=self.bnb_quantized)
        return new


class Linear4bit(torch.nn.Linear):
    """NF4 linear layer; ``forward`` runs the fused decode + tcgen05 GEMM (no bf16 weight in HBM).

    Constructor signature and attributes follow ``bitsandbytes.nn.Linear4bit`` (call site:
    transformers' ``replace_with_bnb_linear``, triggered by load_cullavo.py:73-86).
    """

    def __init__(self, input_features, output_features, bias=True, compute_dtype=None, compress_statistics=True,
                 quant_type="fp4", quant_storage=torch.uint8, device=None):
        super().__init__(input_features, output_features, bias, device)
        self.weight = Params4bit(self.weight.data, requires_grad=False, compress_statistics=compress_statistics,
                                 quant_type=quant_type, quant_storage=quant_storage, module=self)
        self.compute_dtype = compute_dtype
        self.compute_type_is_set = compute_dtype is not None
        self.quant_state = None
        self.quant_storage = quant_storage

    def set_compute_type(self, x):
        if x.dtype in (torch.float32, torch.bfloat16):
            self.compute_dtype = x.dtype
        elif x.dtype == torch.float16:
            warnings.warn("Input type into Linear4bit is torch.float16; the sm_100a kernels compute in bfloat16.")

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        """weight (uint8) + bitsandbytes' quant-state component keys."""
        super()._save_to_state_dict(destination, prefix, keep_vars)
        if getattr(self.weight, "quant_state", None) is not None:
            for k, v in self.weight.quant_state.as_dict(packed=True).items():
                destination[prefix + "weight." + k] = v if keep_vars else v.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        qs_keys = [k for k in state_dict if k.startswith(prefix + "weight.")]
        if qs_keys and (prefix + "weight") in state_dict:
            stats = {k[len(prefix + "weight."):]: state_dict.pop(k) for k in qs_keys}
            data = state_dict.pop(prefix + "weight")
            dev = data.device if data.is_cuda else (self.weight.device if self.weight.is_cuda else data.device)
            self.weight = Params4bit.from_prequantized(data, stats, device=dev, module=self)
            if self.bias is not None and (prefix + "bias") in state_dict:
                with torch.no_grad():
                    self.bias.copy_(state_dict.pop(prefix + "bias"))
            return
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def _quant_state(self):
        qs = getattr(self.weight, "quant_state", None)
        if qs is None:
            if getattr(self, "quant_state", None) is not None:
                # the weight lost its state through a generic .to()/FSDP round trip; recover it
                self.weight.quant_state = self.quant_state
                qs = self.quant_state
            else:
                raise RuntimeError("quantization state not initialized: call .cuda() / .to('cuda') on the module "
                                   "first (weights are quantised to NF4 when they move to the GPU)")
        return qs

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        qs = self._quant_state()
        if not self.compute_type_is_set:
            self.set_compute_type(x)
            self.compute_type_is_set = True
        inp_dtype = x.dtype
        out = matmul_4bit(x, self.weight, qs, bias=self.bias)
        return out.to(inp_dtype)
