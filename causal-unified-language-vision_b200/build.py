"""Builds libb2q.so (all CUDA kernels + the C ABI) for sm_100a with nvcc, in-tree.

    python causal-unified-language-vision_b200/build.py [--force] [--verbose] [--variant NAME -DMACRO=V ...]

``--variant NAME`` builds ``libb2q_NAME.so`` with the extra ``-D`` switches (same-box A/B runs of two builds through
``B2Q_LIB_PATH``; never loaded by default).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb2q.so")
SOURCES = ["nf4_kernels.cu", "qlora_gemm.cu", "optim_kernels.cu", "comm.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC",
]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b2q.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    out_path = OUT if not variant else os.path.join(HERE, f"libb2q_{variant}.so")
    if not variant and not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    build_dir = os.path.join(HERE, "build" if not variant else f"build_{variant}")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc, "-shared", "-o", out_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return out_path


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=_variant,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
