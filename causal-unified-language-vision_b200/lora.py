"""LoRA adapter surface over ``Linear4bit``: what ``model.add_adapter(LoraConfig(...))`` builds.

The reference calls only construction-time PEFT APIs (/root/reference/cullavo/load_cullavo.py:
``LoraConfig`` :94-110 and :24-40, ``add_adapter`` :111-112 and :41-42,
``prepare_model_for_kbit_training`` :91-93) and then uses generic ``nn.Module`` features
(``named_parameters`` with ``'lora' in name`` filters, modeling/BaseModel.py:76-79).  This file
keeps those names, argument meanings and parameter paths
(``<proj>.base_layer.weight``, ``<proj>.lora_A.<adapter>.weight``, ``<proj>.lora_B.<adapter>.weight``)
and routes ``forward`` to the fused sm_100a kernels.
"""
from __future__ import annotations

import math
import re
import warnings
from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Union

import torch
import torch.nn as nn

from .autograd import qlora_linear
from .nn import Linear4bit, Params4bit


@dataclass
class LoraConfig:
    """Subset of ``peft.LoraConfig`` the reference uses (load_cullavo.py:24-40, 94-110)."""

    r: int = 8
    lora_alpha: int = 8
    target_modules: Optional[Union[List[str], str]] = None
    lora_dropout: float = 0.0
    bias: str = "none"
    task_type: Optional[str] = None
    layers_to_transform: Optional[Union[List[int], int]] = None
    layers_pattern: Optional[Union[List[str], str]] = None
    init_lora_weights: bool = True
    modules_to_save: Optional[List[str]] = None
    inference_mode: bool = False
    peft_type: str = field(default="LORA", init=False)

    def __post_init__(self):
        if self.bias != "none":
            raise NotImplementedError("only bias='none' (the reference's setting)")
        if isinstance(self.target_modules, (list, tuple, set)):
            self.target_modules = list(self.target_modules)


class LoraLayer:
    """Duck-type of ``peft.tuners.lora.LoraLayer`` (attribute names are part of the checkpoint/loader contract)."""

    adapter_layer_names = ("lora_A", "lora_B")
    other_param_names = ("r", "lora_alpha", "scaling", "lora_dropout")

    def __init__(self, base_layer: nn.Module, **kwargs) -> None:
        self.base_layer = base_layer
        self.r = {}
        self.lora_alpha = {}
        self.scaling = {}
        self.lora_dropout = nn.ModuleDict({})
        self.lora_A = nn.ModuleDict({})
        self.lora_B = nn.ModuleDict({})
        self._disable_adapters = False
        self.merged_adapters = []
        self._active_adapter = "default"
        self.in_features = base_layer.in_features
        self.out_features = base_layer.out_features
        self.kwargs = kwargs

    # -- PEFT API --------------------------------------------------------------------------
    def update_layer(self, adapter_name, r, lora_alpha, lora_dropout, init_lora_weights=True, use_rslora=False,
                     use_dora=False):
        if r <= 0:
            raise ValueError(f"`r` should be a positive integer value but the value passed is {r}")
        if use_dora:
            raise NotImplementedError("DoRA is not part of the reference path")
        self.r[adapter_name] = r
        self.lora_alpha[adapter_name] = lora_alpha
        self.lora_dropout[adapter_name] = nn.Dropout(p=lora_dropout) if lora_dropout > 0.0 else nn.Identity()
        self.lora_A[adapter_name] = nn.Linear(self.in_features, r, bias=False)
        self.lora_B[adapter_name] = nn.Linear(r, self.out_features, bias=False)
        self.scaling[adapter_name] = lora_alpha / math.sqrt(r) if use_rslora else lora_alpha / r
        if init_lora_weights:
            self.reset_lora_parameters(adapter_name, init_lora_weights)
        # adapters live on the device of the (quantised) base weight, created in fp32 like PEFT does
        weight = getattr(self.base_layer, "weight", None)
        if weight is not None and weight.device.type != "meta":
            self.lora_A[adapter_name].to(weight.device)
            self.lora_B[adapter_name].to(weight.device)
        self.set_adapter(self.active_adapters)

    def reset_lora_parameters(self, adapter_name, init_lora_weights=True):
        if init_lora_weights is False:
            return
        if adapter_name in self.lora_A.keys():
            nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B[adapter_name].weight)

    @property
    def merged(self) -> bool:
        return bool(self.merged_adapters)

    @property
    def disable_adapters(self) -> bool:
        return self._disable_adapters

    @property
    def active_adapter(self):
        return self._active_adapter

    @property
    def active_adapters(self):
        if isinstance(self._active_adapter, str):
            return [self._active_adapter]
        return self._active_adapter

    def set_adapter(self, adapter_names) -> None:
        """Activate adapters; only active adapters' A/B require grad (PEFT semantics)."""
        if isinstance(adapter_names, str):
            adapter_names = [adapter_names]
        for layer_name in self.adapter_layer_names:
            module_dict = getattr(self, layer_name)
            for key, layer in module_dict.items():
                layer.requires_grad_(key in adapter_names)
        self._active_adapter = adapter_names

    def enable_adapters(self, enabled: bool) -> None:
        if enabled:
            self.set_adapter(self.active_adapters)
            self._disable_adapters = False
        else:
            for layer_name in self.adapter_layer_names:
                getattr(self, layer_name).requires_grad_(False)
            self._disable_adapters = True

    def merge(self, *args, **kwargs):
        raise NotImplementedError("merging into an NF4 weight re-quantises it; not part of the training path")


class LoraLinear4bit(nn.Module, LoraLayer):
    """``peft.tuners.lora.bnb.Linear4bit`` equivalent; forward = ONE fused QLoRA linear.

    Reference semantics: ``result = base_layer(x).clone();
    result += lora_B(lora_A(dropout(x))) * scaling`` for the one active adapter.
    """

    def __init__(self, base_layer: Linear4bit, adapter_name: str, r: int = 0, lora_alpha: int = 1,
                 lora_dropout: float = 0.0, init_lora_weights: bool = True, **kwargs) -> None:
        super().__init__()
        LoraLayer.__init__(self, base_layer)
        self._active_adapter = adapter_name
        self.update_layer(adapter_name, r, lora_alpha, lora_dropout, init_lora_weights)
        self._grad_sinks = {}  # adapter -> parallel.GradSync (owns the flat gradient buckets)

    # PEFT exposes these on the wrapper as well
    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        active = [a for a in self.active_adapters if a in self.lora_A.keys()]
        if self.disable_adapters or not active:
            return self.base_layer(x, *args, **kwargs)
        if len(active) > 1:
            raise NotImplementedError("one active adapter at a time (the reference activates exactly one: "
                                      "HF add_adapter ends with set_adapter(name))")
        name = active[0]
        base = self.base_layer
        qs = base._quant_state()
        A = self.lora_A[name].weight
        B = self.lora_B[name].weight
        drop = self.lora_dropout[name]
        p = float(drop.p) if (isinstance(drop, nn.Dropout) and self.training) else 0.0
        # seed drawn from torch's CPU generator: torch.utils.checkpoint restores that state on
        # recompute, so the re-run forward regenerates the same mask
        seed = int(torch.randint(0, 2**62, (1,)).item()) if p > 0.0 else 0
        sink = None
        if name in self._grad_sinks and torch.is_grad_enabled():
            sink = self._grad_sinks[name].sink_for(self)  # parallel.GradSync: grads go straight into its buckets
        inp_dtype = x.dtype
        y = qlora_linear(x, base.weight.data, qs, A, B, self.scaling[name], p, seed, sink)
        if base.bias is not None:
            y = y + base.bias.to(y.dtype)
        return y.to(inp_dtype)

    def __repr__(self) -> str:
        return "lora." + super().__repr__()


# ----------------------------------------------------------------------- model surgery ----
def _get_submodules(model: nn.Module, key: str):
    parent_name, _, target_name = key.rpartition(".")
    parent = model.get_submodule(parent_name) if parent_name else model
    return parent, model.get_submodule(key), target_name


def replace_with_4bit_linear(model: nn.Module, modules_to_not_convert: Optional[Iterable[str]] = None,
                             compute_dtype=torch.bfloat16, compress_statistics: bool = True, quant_type: str = "nf4",
                             quant_storage=torch.uint8) -> nn.Module:
    """Swap every ``nn.Linear`` (except names in ``modules_to_not_convert``) for ``Linear4bit``.

    Stand-in for transformers' ``replace_with_bnb_linear`` driven by the ``BitsAndBytesConfig`` at
    load_cullavo.py:73-82 (``llm_int8_skip_modules=['multi_modal_projector','lm_head']``).  The
    new modules keep the fp weight until ``.cuda()`` / ``.to('cuda')`` quantises it.
    """
    skip = list(modules_to_not_convert or [])
    for name, module in list(model.named_modules()):
        if not isinstance(module, nn.Linear) or isinstance(module, Linear4bit):
            continue
        if any((s + "." in name + ".") or (name == s) or name.endswith("." + s) for s in skip):
            continue
        parent, target, target_name = _get_submodules(model, name)
        new = Linear4bit(target.in_features, target.out_features, bias=target.bias is not None,
                         compute_dtype=compute_dtype, compress_statistics=compress_statistics, quant_type=quant_type,
                         quant_storage=quant_storage, device="meta")
        was_cuda = target.weight.is_cuda
        new.weight = Params4bit(target.weight.data.detach().to("cpu") if was_cuda else target.weight.data.detach(),
                                requires_grad=False, compress_statistics=compress_statistics, quant_type=quant_type,
                                quant_storage=quant_storage, module=new)
        if target.bias is not None:
            new.bias = nn.Parameter(target.bias.data.detach().clone(), requires_grad=False)
        new.requires_grad_(False)
        setattr(parent, target_name, new)
        if was_cuda:
            new.to(target.weight.device)
    return model


def _matches(key: str, config: LoraConfig) -> bool:
    tm = config.target_modules
    if tm is None:
        return False
    if isinstance(tm, str):
        hit = re.fullmatch(tm, key) is not None
    else:
        hit = any(key == t or key.endswith("." + t) for t in tm)
    if hit and config.layers_to_transform is not None:
        idx = [config.layers_to_transform] if isinstance(config.layers_to_transform, int) else list(
            config.layers_to_transform)
        m = re.match(r".*\.[^.]*\.(\d+)\.", "." + key) if not config.layers_pattern else None
        if config.layers_pattern:
            pats = [config.layers_pattern] if isinstance(config.layers_pattern, str) else config.layers_pattern
            for pat in pats:
                m = re.match(rf".*\.{pat}\.(\d+)\.", "." + key)
                if m:
                    break
        hit = m is not None and int(m.group(1)) in idx
    return hit


def add_adapter(model: nn.Module, config: LoraConfig, adapter_name: str = "default") -> nn.Module:
    """``model.add_adapter(config, adapter_name=...)`` (HF PEFT integration) for ``Linear4bit`` targets.

    Like the HF method it injects the adapter, freezes everything that is not a LoRA weight of
    an active adapter and ends with ``set_adapter(adapter_name)`` -- exactly one adapter is
    active afterwards.
    """
    found = False
    for key, module in list(model.named_modules()):
        if isinstance(module, LoraLinear4bit):
            if _matches(key, config):
                module.update_layer(adapter_name, config.r, config.lora_alpha, config.lora_dropout,
                                    config.init_lora_weights)
                found = True
            continue
        if not isinstance(module, Linear4bit) or not _matches(key, config):
            continue
        if key.endswith(".base_layer"):
            continue
        parent, target, target_name = _get_submodules(model, key)
        new = LoraLinear4bit(target, adapter_name, r=config.r, lora_alpha=config.lora_alpha,
                             lora_dropout=config.lora_dropout, init_lora_weights=config.init_lora_weights)
        setattr(parent, target_name, new)
        found = True
    if not found:
        raise ValueError(f"Target modules {config.target_modules} not found in the base model.")
    for n, p_ in model.named_parameters():
        if "lora_" not in n:
            p_.requires_grad = False
    set_adapter(model, adapter_name)
    if not hasattr(model, "peft_config"):
        model.peft_config = {}
    model.peft_config[adapter_name] = config
    model._hf_peft_config_loaded = True
    return model


def set_adapter(model: nn.Module, adapter_name: Union[str, List[str]]) -> None:
    for m in model.modules():
        if isinstance(m, LoraLayer):
            m.set_adapter(adapter_name)
            m._disable_adapters = False


def prepare_model_for_kbit_training(model: nn.Module, use_gradient_checkpointing: bool = True,
                                    gradient_checkpointing_kwargs: Optional[dict] = None) -> nn.Module:
    """``peft.prepare_model_for_kbit_training`` (load_cullavo.py:91-93): freeze the base, upcast the
    remaining half-precision parameters to fp32, switch on (non-reentrant) gradient checkpointing."""
    for _, param in model.named_parameters():
        param.requires_grad = False
    for param in model.parameters():
        if param.dtype in (torch.float16, torch.bfloat16) and not isinstance(param, Params4bit):
            param.data = param.data.to(torch.float32)
    if use_gradient_checkpointing:
        if hasattr(model, "enable_input_require_grads"):
            model.enable_input_require_grads()
        if hasattr(model, "gradient_checkpointing_enable"):
            model.gradient_checkpointing_enable(gradient_checkpointing_kwargs=gradient_checkpointing_kwargs or {})
        else:
            warnings.warn("model has no gradient_checkpointing_enable(); wrap layers with torch.utils.checkpoint")
    return model


def find_all_linear_names(model: nn.Module) -> List[str]:
    """The reference's own helper (load_cullavo.py:8-20), kept to show the module surface satisfies it."""
    names = set()
    for name, module in model.named_modules():
        if isinstance(module, nn.Linear):
            parts = name.split(".")
            if "out_proj" in parts[-1]:
                continue
            names.add(parts[0] if len(parts) == 1 else parts[-1])
    names.discard("lm_head")
    return sorted(names)


# ----------------------------------------------------------------- adapter checkpoints ----
def get_adapter_state_dict(model: nn.Module, adapter_name: str = "default") -> dict:
    """PEFT on-disk naming: ``base_model.model.<path>.lora_A.weight`` (adapter name stripped)."""
    out = {}
    for k, v in model.state_dict().items():
        if "lora_" in k and f".{adapter_name}." in k:
            out["base_model.model." + k.replace(f".{adapter_name}.", ".")] = v.detach()
    return out


def set_adapter_state_dict(model: nn.Module, state: dict, adapter_name: str = "default") -> None:
    own = dict(model.named_parameters())
    for k, v in state.items():
        k = k[len("base_model.model."):] if k.startswith("base_model.model.") else k
        name = re.sub(r"\.(lora_[AB])\.weight$", rf".\1.{adapter_name}.weight", k)
        if name not in own:
            raise KeyError(f"adapter tensor {k} has no parameter {name} in the model")
        with torch.no_grad():
            own[name].copy_(v.to(own[name].dtype))


def save_adapter(model: nn.Module, path: str, adapter_name: str = "default") -> None:
    from safetensors.torch import save_file

    save_file({k: v.contiguous().cpu() for k, v in get_adapter_state_dict(model, adapter_name).items()}, path)


def load_adapter(model: nn.Module, path: str, adapter_name: str = "default") -> None:
    from safetensors.torch import load_file

    set_adapter_state_dict(model, load_file(path), adapter_name)
