"""examples/c_host_example.c: the hot path driven from plain C through the C ABI (no Python, no torch in the process)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "causal-unified-language-vision_b200")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_example(out_dir: str) -> str:
    exe = os.path.join(out_dir, "c_host_example")
    cmd = ["gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
           os.path.join(ROOT, "examples", "c_host_example.c"), "-L" + PKG, "-lb2q", "-L" + os.path.join(CUDA, "lib64"),
           "-lcudart", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not on PATH")
def test_c_host_example_compiles_and_links_as_plain_c(lib_built, tmp_path):
    """include/b2q.h is valid C11 and libb2q.so satisfies every symbol a C host uses (no GPU needed to link)."""
    build_example(str(tmp_path))


@pytest.mark.gpu
def test_c_host_example_runs(lib_built, tmp_path):
    exe = build_example(str(tmp_path))
    env = dict(os.environ, LD_LIBRARY_PATH=PKG + ":" + os.path.join(CUDA, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "C host example OK" in r.stdout, r.stdout + r.stderr
