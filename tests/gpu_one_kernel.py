"""Runs one main kernel a few times (for ncu captures).  usage: gpu_one_kernel.py fwd|dx VARIANT [M N K] [lora]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import b200qlora as q  # noqa: E402

F = q.functional
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 5
M, N, K = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (16384, 4096, 4096)
lora = len(sys.argv) > 6 and sys.argv[6] == "lora"
dev = torch.device("cuda:0")
torch.manual_seed(0)
packed, qs = F.quantize_4bit(torch.randn(N, K, device=dev) * 0.02, compress_statistics=True)
x = torch.randn(M, K, device=dev).bfloat16()
dy = torch.randn(M, N, device=dev).bfloat16()
A = (torch.randn(64, K, device=dev) * 0.01).bfloat16()
B = (torch.randn(N, 64, device=dev) * 0.02).bfloat16()
us = du = None
if lora:
    _, us = F.lora_down(x, A, 0.25)
    du = F.lora_bwd_du(dy, B, 0.25)
F.set_variant(variant, variant)
torch.cuda.synchronize()
for _ in range(4):
    if which == "fwd":
        y = F.qlora_fwd(x, packed, qs, us, B if lora else None)
    else:
        y = F.qlora_bwd_dx(dy, packed, qs, du, A if lora else None)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
