"""Dry run of bench.py's control flow on a machine without a GPU (development tool, not a test of the hot path).

Everything that touches the device is replaced by a stub: the stack is a fake whose steps do nothing, CUDA events return
a fixed time, "cuda" tensors are CPU tensors.  What is exercised for real: argument handling, the value / e2e / optimizer
phases' bookkeeping, the watchdog wiring, and the JSON line (keys, types).  Usage:

    python tests/dev_bench_dryrun.py            # N = 1 flow
    python tests/dev_bench_dryrun.py --world 2  # N > 1 flow with a fake process group (single process)
    python tests/dev_bench_dryrun.py [--world 2] --hang modules   # module-surface step blocks: the watchdog prints the line
    python tests/dev_bench_dryrun.py --hang raise                 # the step dies with a CUDA-like error: same line, rc 5
"""
import argparse
import contextlib
import importlib
import json
import os
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--hang", default="", choices=["", "modules", "raise"])
ap.add_argument("--e2e-timeout", type=float, default=3.0)
opts = ap.parse_args()

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

CPU = torch.device("cpu")


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class FakeStream:
    cuda_stream = 0

    def synchronize(self):
        pass

    def wait_event(self, ev):
        pass


real_device = torch.device


def fake_device(*a, **k):
    if a and a[0] == "cuda":
        return CPU
    return real_device(*a, **k)


torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.Event = FakeEvent
torch.cuda.current_stream = lambda *a: FakeStream()
torch.cuda.Stream = lambda *a, **k: FakeStream()
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.Tensor.record_stream = lambda self, s: None
torch.device = fake_device
torch.Tensor.pin_memory = lambda self, *a, **k: self

if opts.world > 1:
    os.environ.update(WORLD_SIZE=str(opts.world), RANK="0", LOCAL_RANK="0")
    dist.init_process_group = lambda *a, **k: None
    dist.barrier = lambda *a, **k: None
    dist.all_reduce = lambda t, *a, **k: None
    dist.destroy_process_group = lambda *a, **k: None
    dist.is_initialized = lambda: False

q = importlib.import_module("b200qlora")
stackmod = importlib.import_module("causal-unified-language-vision_b200.stack")
F = q.functional
counter = [0]
F.launch_count = lambda: counter[0]


def fake_op(*a):
    counter[0] += 1
    return None


F.qlora_fwd = fake_op
F.qlora_bwd_dx = fake_op


class FakeQS:
    shape = (64, 128)


class FakeSync:
    def flat_grads(self):
        return [torch.ones(8)]


class FakeStack:
    def __init__(self, layers, shapes, M, r=64, dropout=0.05, device=None, seed=0):
        self.shapes, self.M = shapes, 16
        self.inputs = {k: torch.zeros(16, 8, dtype=torch.bfloat16) for _, _, k in shapes}
        self.grads_out = {n: torch.zeros(16, 8, dtype=torch.bfloat16) for _, n, _ in shapes}
        self.sync = FakeSync()
        self._fpt = stackmod.stack_flops_per_token(shapes, layers, r)

    def flops_per_token(self):
        return self._fpt

    def step_direct(self, recompute=False, inputs=None, grads_out=None):
        x = torch.zeros(128, 128, dtype=torch.bfloat16)
        F.qlora_fwd(x, None, FakeQS(), None, None)
        F.qlora_bwd_dx(x, None, FakeQS(), None, None)
        time.sleep(0.01)

    def step_modules(self, inputs=None, grads_out=None, trace=None, interleaved=False):
        if opts.hang == "modules":
            time.sleep(3600)
        if opts.hang == "raise":
            raise RuntimeError("CUDA error: unspecified launch failure (dry-run stand-in)")
        assert set(inputs) == {k for _, _, k in self.shapes} and set(grads_out) == {n for _, n, _ in self.shapes}
        time.sleep(0.01)
        return self.grad_sqnorm()

    def grad_sqnorm(self):
        return torch.tensor(8.0)


import inspect  # noqa: E402

for name in ("step_direct", "step_modules", "grad_sqnorm", "flops_per_token"):   # the fake mirrors the real signatures
    real_sig = inspect.signature(getattr(stackmod.QLoRALinearStack, name))
    fake_sig = inspect.signature(getattr(FakeStack, name))
    assert list(real_sig.parameters) == list(fake_sig.parameters), (name, real_sig, fake_sig)
stackmod.QLoRALinearStack = FakeStack

sys.argv = ["bench.py", "--gpus", str(opts.world), "--steps", "2", "--warmup", "3", "--no-cpu", "--no-opt",
            "--e2e-timeout", str(opts.e2e_timeout), "--global-timeout", "30"]
bench = importlib.import_module("bench")

import io  # noqa: E402
import contextlib  # noqa: E402

if opts.hang:
    # the watchdog ends the process with os._exit(5) after printing the line: run and let it
    bench.main()
    raise SystemExit("watchdog did not fire")
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
line = [l for l in buf.getvalue().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
    assert key in d, key
assert d["n_gpus"] == opts.world and d["e2e"] is not None and d["e2e"]["h2d_bytes_per_step"] > 0
assert d["e2e"]["api"].startswith("LoraLinear4bit.forward")
print("dry run OK:", json.dumps({k: d[k] for k in ("n_gpus", "value", "e2e", "gpu_launches")})[:600])
