"""Runs each HBM-bound LoRA kernel a few times at BASELINE shapes (for ncu captures of the non-GEMM rows)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import b200qlora as q  # noqa: E402

F = q.functional
dev = torch.device("cuda:0")
M, N, K, r = 16384, 4096, 4096, 64
torch.manual_seed(0)
packed, qs = F.quantize_4bit(torch.randn(N, K, device=dev) * 0.02, compress_statistics=True)
x = torch.randn(M, K, device=dev).bfloat16()
dy = torch.randn(M, N, device=dev).bfloat16()
A = (torch.randn(r, K, device=dev) * 0.01).bfloat16()
B = (torch.randn(N, r, device=dev) * 0.02).bfloat16()
dA, dB = torch.zeros_like(A), torch.zeros_like(B)
for _ in range(3):
    u, us = F.lora_down(x, A, 0.25, 5, 0.05)
    du = F.lora_bwd_du(dy, B, 0.25, 0.05)
    dx = F.qlora_bwd_dx(dy, packed, qs, du, A, 5, 0.05)
    F.lora_grads(dy, x, u, du, 0.25, dA, dB, seed=5, p=0.05)
torch.cuda.synchronize()
print("ok")
