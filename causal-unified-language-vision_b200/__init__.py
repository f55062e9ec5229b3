"""B200-native QLoRA linear hot path (NF4 Linear4bit + LoRA, forward and backward).

Directory name follows the build contract (``causal-unified-language-vision_b200``); because of the
hyphens it is imported through the ``b200qlora`` alias module at the repo root
(``import b200qlora``) or ``importlib.import_module("causal-unified-language-vision_b200")``.

    functional  tensor-level ops over the C ABI (quantize_4bit, dequantize_4bit, qlora_fwd, ...)
    nn          Params4bit / Linear4bit          (bitsandbytes module surface)
    lora        LoraConfig / LoraLinear4bit / add_adapter / prepare_model_for_kbit_training  (PEFT surface)
    autograd    MatMul4Bit / QLoRALinear autograd Functions
    parallel    GradSync: flat LoRA-gradient buckets + overlapped NCCL all-reduce
    optim       FusedLoraAdamW: AdamW + global-norm clip over the flat LoRA buckets
    stack       the linear-stack workload used by bench.py
"""
from . import _lib, functional  # noqa: F401
from . import autograd, nn, lora, parallel, stack, optim  # noqa: F401,E402
from .functional import QuantState, quantize_4bit, dequantize_4bit  # noqa: F401
from .nn import Linear4bit, Params4bit  # noqa: F401
from .lora import LoraConfig, LoraLinear4bit, add_adapter, prepare_model_for_kbit_training  # noqa: F401

__all__ = ["functional", "nn", "lora", "autograd", "parallel", "stack", "QuantState", "quantize_4bit",
           "dequantize_4bit", "Linear4bit", "Params4bit", "LoraConfig", "LoraLinear4bit", "add_adapter",
           "prepare_model_for_kbit_training"]
