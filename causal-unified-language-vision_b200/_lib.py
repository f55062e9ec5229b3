"""ctypes binding of libb2q.so (the C ABI declared in include/b2q.h).

There is no CPU implementation and no fallback: if the shared library is missing the import
of any op raises, and every call checks the return code and raises ``RuntimeError``.
Mirrors how bitsandbytes itself is layered (Python -> ctypes -> ``extern "C"`` functions
taking raw pointers + ``cudaStream_t``), which is the interface the reference's dependencies
bind.
"""
from __future__ import annotations

import ctypes as ct
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2Q_LIB_PATH: development hook for same-box A/B runs of two builds of the library (never a fallback)
LIB_PATH = os.environ.get("B2Q_LIB_PATH") or os.path.join(_HERE, "libb2q.so")

c_void_p = ct.c_void_p
c_int = ct.c_int
c_i64 = ct.c_int64
c_u64 = ct.c_uint64
c_float = ct.c_float
c_size_t = ct.c_size_t


class NF4Weight(ct.Structure):
    """``struct b2q_nf4_weight`` of include/b2q.h."""

    _fields_ = [
        ("packed", c_void_p),
        ("absmax", c_void_p),
        ("absmax_q", c_void_p),
        ("absmax2", c_void_p),
        ("code256", c_void_p),
        ("offset", c_float),
        ("code16", c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol include/b2q.h declares
SIGNATURES = {
    "b2q_version": (c_int, []),
    "b2q_error_string": (ct.c_char_p, [c_int]),
    "b2q_last_error_detail": (ct.c_char_p, []),
    "b2q_nf4_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_i64,
                               c_int, c_int, c_void_p]),
    "b2q_nf4_quantize": (c_int, [c_void_p, c_int, c_i64, c_void_p, c_void_p, c_void_p]),
    "b2q_absmax_double_quant": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2q_gemv_4bit": (c_int, [c_void_p, ct.POINTER(NF4Weight), c_void_p, c_int, c_int, c_int, c_void_p]),
    "b2q_dropout_mask": (c_int, [c_void_p, c_i64, c_u64, c_float, c_void_p]),
    "b2q_dropout_apply": (c_int, [c_void_p, c_void_p, c_i64, c_u64, c_float, c_void_p]),
    "b2q_dropout_bwd_add": (c_int, [c_void_p, c_void_p, c_i64, c_u64, c_float, c_void_p]),
    "b2q_lora_down": (c_int, [c_void_p, c_void_p, c_float, c_u64, c_float, c_void_p, c_void_p, c_int, c_int, c_int,
                              c_void_p]),
    "b2q_qlora_fwd": (c_int, [c_void_p, ct.POINTER(NF4Weight), c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                              c_int, c_void_p]),
    "b2q_lora_bwd_du": (c_int, [c_void_p, c_void_p, c_float, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b2q_qlora_bwd_dx": (c_int, [c_void_p, ct.POINTER(NF4Weight), c_void_p, c_void_p, c_u64, c_float, c_void_p, c_int,
                                 c_int, c_int, c_int, c_void_p]),
    "b2q_lora_grads_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "b2q_lora_grads": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_u64, c_float, c_void_p, c_void_p,
                               c_int, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_void_p]),
    "b2q_gemm_bf16": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b2q_reduce_partials": (c_int, [c_void_p, c_int, c_i64, c_float, c_void_p, c_int, c_void_p]),
    "b2q_sqnorm_blocks": (c_int, [c_i64]),
    "b2q_sqnorm_partials": (c_int, [c_void_p, c_i64, c_void_p, c_void_p]),
    "b2q_adamw_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_float, c_float, c_float, c_float,
                               c_float, c_i64, c_void_p, c_int, c_float, c_void_p]),
    "b2q_comm_nccl_version": (c_int, []),
    "b2q_comm_unique_id": (c_int, [c_void_p, c_size_t]),
    "b2q_comm_init": (c_int, [ct.POINTER(c_void_p), c_void_p, c_size_t, c_int, c_int]),
    "b2q_comm_allreduce_bucket": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p]),
    "b2q_comm_wait": (c_int, [c_void_p, c_void_p]),
    "b2q_comm_destroy": (c_int, [c_void_p]),
    "b2q_set_variant": (c_int, [c_int, c_int]),
    "b2q_debug_set_trace": (c_int, [c_void_p, c_int]),
    "b2q_debug_set_prefetch": (c_int, [c_int]),
    "b2q_debug_stall_count": (c_int, []),
    "b2q_debug_stall_report": (c_int, [ct.c_char_p, c_size_t]),
    "b2q_debug_stall_selftest": (c_int, [c_void_p]),
    "b2q_launch_count": (c_u64, []),
}

_lib = None


def load() -> ct.CDLL:
    """Load libb2q.so (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA library has not been built (run `python -c 'import "
            "__graft_entry__ as g; g.build()'` at the repo root).  There is no CPU fallback."
        )
    lib = ct.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def stall_report() -> str:
    """Text of the kernels' stall-guard records ('' when there are none).  Works after the CUDA context has been lost:
    the records live in host memory (include/b2q.h, "Stall guard")."""
    lib = load()
    if lib.b2q_debug_stall_count() == 0:
        return ""
    buf = ct.create_string_buffer(1 << 16)
    lib.b2q_debug_stall_report(buf, len(buf))
    return buf.value.decode(errors="replace")


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().b2q_error_string(code)
        detail = load().b2q_last_error_detail() if code < 0 else b""
        stalls = stall_report() if code > 0 else ""
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else code} (code {code})"
                           + (f" [{detail.decode()}]" if detail else "") + (f"\n{stalls}" if stalls else ""))
