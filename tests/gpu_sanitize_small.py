"""Small end-to-end pass of every kernel family for compute-sanitizer (memcheck): tiny shapes, all code paths
(nested absmax, LoRA tail, dropout, ragged M, GEMV, optimizer).  usage: compute-sanitizer --tool memcheck python tests/gpu_sanitize_small.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import b200qlora as q  # noqa: E402

F = q.functional
dev = torch.device("cuda:0")
torch.manual_seed(0)
M, N, K, r = 200, 512, 768, 64
packed, qs = F.quantize_4bit(torch.randn(N, K, device=dev) * 0.02, compress_statistics=True)
x = torch.randn(M, K, device=dev).bfloat16()
dy = torch.randn(M, N, device=dev).bfloat16()
A = (torch.randn(r, K, device=dev) * 0.01).bfloat16()
B = (torch.randn(N, r, device=dev) * 0.02).bfloat16()
for p in (0.0, 0.05):
    u, us = F.lora_down(x, A, 0.25, 7, p)
    y = F.qlora_fwd(x, packed, qs, us, B)
    du = F.lora_bwd_du(dy, B, 0.25, p)
    dx = F.qlora_bwd_dx(dy, packed, qs, du, A, 7, p)
    dA, dB = torch.zeros_like(A), torch.zeros_like(B)
    F.lora_grads(dy, x, u, du, 0.25, dA, dB, seed=7, p=p)
w = F.dequantize_4bit(packed, qs)
g = F.gemv_4bit(x[:3].contiguous(), packed, qs)
stackmod = importlib.import_module("causal-unified-language-vision_b200.stack")
optim = importlib.import_module("causal-unified-language-vision_b200.optim")
st = stackmod.QLoRALinearStack(1, [("q_proj", 256, 256), ("up_proj", 512, 256)], 96, r=64, dropout=0.05, device=dev)
st.step_direct()
st.step_modules()
o = optim.FusedLoraAdamW(st.sync, lr=1e-3)
o.clip_grad_norm_(1.0)
o.step()
torch.cuda.synchronize()
print("sanitize pass ok", float(y.float().abs().mean()), float(dx.float().abs().mean()), float(g.float().abs().mean()))
