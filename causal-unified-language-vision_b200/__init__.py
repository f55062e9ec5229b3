"""B200-native QLoRA linear hot path (NF4 Linear4bit + LoRA, forward and backward).

Directory name follows the build contract (``causal-unified-language-vision_b200``); because of the
hyphens it is imported through the ``b200qlora`` alias module at the repo root
(``import b200qlora``) or ``importlib.import_module("causal-unified-language-vision_b200")``.
"""
from . import _lib, functional  # noqa: F401
from .functional import QuantState, quantize_4bit, dequantize_4bit  # noqa: F401

__all__ = ["functional", "QuantState", "quantize_4bit", "dequantize_4bit"]
