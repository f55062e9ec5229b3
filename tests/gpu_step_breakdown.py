"""In-step time per kernel class (CUDA events around every C-ABI call of one full linear-stack step)."""
import collections
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import b200qlora as q  # noqa: E402

F = q.functional
stackmod = importlib.import_module("causal-unified-language-vision_b200.stack")
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
st = stackmod.QLoRALinearStack(layers, stackmod.MISTRAL_LITERAL, 16384, r=64, dropout=0.05, device=dev, seed=0)
events = []
names = ["lora_down", "qlora_fwd", "lora_bwd_du", "qlora_bwd_dx", "lora_grads"]
orig = {n: getattr(F, n) for n in names}
on = [False]


def wrap(n):
    def f(*a, **k):
        if not on[0]:
            return orig[n](*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig[n](*a, **k)
        e1.record()
        events.append((n, e0, e1))
        return out
    return f


for n in names:
    setattr(F, n, wrap(n))
for _ in range(3):
    st.step_direct()
torch.cuda.synchronize()
on[0] = True
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
steps = 3
for _ in range(steps):
    st.step_direct()
t1.record()
torch.cuda.synchronize()
tot = t0.elapsed_time(t1) / steps
agg = collections.OrderedDict()
for n, a, b in events:
    agg[n] = agg.get(n, 0.0) + a.elapsed_time(b) / steps
print(f"step {tot:.1f} ms")
s = 0.0
for n, v in agg.items():
    print(f"  {n:14s} {v:8.2f} ms  {v / tot:6.1%}")
    s += v
print(f"  {'(gaps/other)':14s} {tot - s:8.2f} ms  {(tot - s) / tot:6.1%}")
