"""The C-ABI library builds, loads, and exports every symbol include/b2q.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b2q.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2q_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = _declared_symbols()
    for must in ("b2q_nf4_decode", "b2q_nf4_quantize", "b2q_qlora_fwd", "b2q_qlora_bwd_dx", "b2q_lora_grads",
                 "b2q_lora_down", "b2q_lora_bwd_du"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/b2q.h but not exported by libb2q.so"


def test_python_binding_covers_header(lib_built):
    import b200qlora as q

    lib = q._lib.load()
    assert lib.b2q_version() == 100
    assert set(_declared_symbols()) == set(q._lib.SIGNATURES), "ctypes signatures out of sync with include/b2q.h"
    assert b"shape" in lib.b2q_error_string(-1)


def test_sass_is_blackwell_native(lib_built):
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG must be in the built library."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        import pytest

        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", lib_built], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "PRMT"):
        assert mnemonic in sass, mnemonic
    # no legacy mma.sync path in any GEMM: the only kernel allowed to hold HMMA is the opt-in GEMV experiment, whose
    # 8-token B fragment is the natural mma.sync shape and whose bound is HBM, not the tensor pipe
    functions = sass.split("Function : ")[1:]
    assert functions
    import re

    legacy = re.compile(r"(?<![A-Z])HMMA\.")   # mma.sync's HMMA.16816, not tcgen05's UTCHMMA
    with_hmma = {f.split()[0] for f in functions if legacy.search(f)}
    assert all("nf4_gemv_mma_kernel" in name for name in with_hmma), with_hmma


def test_comm_entry_points_without_a_gpu(lib_built):
    """b2q_comm_*: argument checking and NCCL resolution (dlopen) work on a machine without a GPU; no collective is run."""
    import ctypes as ct

    import b200qlora as q

    lib = q._lib.load()
    assert lib.b2q_comm_unique_id(None, 0) == -2 and lib.b2q_comm_wait(None, None) == -2
    assert lib.b2q_comm_allreduce_bucket(None, None, 0, 0, 0, None) == -2
    assert lib.b2q_comm_destroy(None) == 0
    comm = ct.c_void_p()
    assert lib.b2q_comm_init(ct.byref(comm), None, 0, 2, 0) == -2
    v = lib.b2q_comm_nccl_version()
    assert v == -5 or v >= 21800, v          # B2Q_ERR_COMM if no libnccl.so.2 can be found, else NCCL's version code
    if v > 0:
        import torch  # noqa: F401 -- torch is loaded, so dlopen must have returned torch's own NCCL

        assert v == torch.cuda.nccl.version()[0] * 10000 + torch.cuda.nccl.version()[1] * 100 + torch.cuda.nccl.version()[2]
        assert b"NCCL" in lib.b2q_error_string(-5)
