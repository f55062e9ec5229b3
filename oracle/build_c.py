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


if __name__ == "__main__":
    print(build(True))
