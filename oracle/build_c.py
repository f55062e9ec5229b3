"""Compiles oracle/nf4_ref.c into oracle/_build/libnf4ref.so (test infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libnf4ref.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "nf4_ref.c")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-fno-fast-math", "-ffp-contract=off", "-shared", "-fPIC", "-o", OUT, src, "-lm"],
                   check=True)
    return OUT


DROPOUT_REF = os.path.join(HERE, "_build", "dropout_host_ref")


def build_dropout_ref(force: bool = False) -> str:
    """Host-only compilation (nvcc, no GPU needed) of the product's dropout_keep(); pins oracle/dropout.py."""
    src = os.path.join(HERE, "dropout_host_ref.cu")
    hdr = os.path.join(HERE, "..", "causal-unified-language-vision_b200", "csrc", "b2q_internal.h")
    if not force and os.path.exists(DROPOUT_REF) and os.path.getmtime(DROPOUT_REF) >= max(os.path.getmtime(src),
                                                                                          os.path.getmtime(hdr)):
        return DROPOUT_REF
    os.makedirs(os.path.dirname(DROPOUT_REF), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", DROPOUT_REF, src], check=True)
    return DROPOUT_REF


if __name__ == "__main__":
    print(build(True))
    print(build_dropout_ref(True))
