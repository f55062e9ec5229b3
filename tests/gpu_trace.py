"""Phase trace of one main-kernel launch (b2q_debug_set_trace): where do the cycles of a tile go?
usage: gpu_trace.py fwd|dx VARIANT [M N K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as ct  # noqa: E402

import torch  # noqa: E402

import b200qlora as q  # noqa: E402

F = q.functional
lib = q._lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 4
M, N, K = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (16384, 4096, 4096)
dev = torch.device("cuda:0")
torch.manual_seed(0)
packed, qs = F.quantize_4bit(torch.randn(N, K, device=dev) * 0.02, compress_statistics=True)
x = torch.randn(M, K, device=dev).bfloat16()
dy = torch.randn(M, N, device=dev).bfloat16()
A = (torch.randn(64, K, device=dev) * 0.01).bfloat16()
B = (torch.randn(N, 64, device=dev) * 0.02).bfloat16()
_, us = F.lora_down(x, A, 0.25)
du = F.lora_bwd_du(dy, B, 0.25)
F.set_variant(variant, variant)
lib.b2q_debug_set_prefetch(int(os.environ.get("B2Q_PF", "0")))
run = (lambda: F.qlora_fwd(x, packed, qs, us, B)) if which == "fwd" else (lambda: F.qlora_bwd_dx(dy, packed, qs, du, A))
for _ in range(20):
    run()
torch.cuda.synchronize()
T = 32
buf = torch.zeros(148, T, 8, dtype=torch.int64, device=dev)
lib.b2q_debug_set_trace(ct.c_void_p(buf.data_ptr()), T)
run()
torch.cuda.synchronize()
lib.b2q_debug_set_trace(None, 0)
t = buf.cpu()
for cta in (0, 1, 74, 146):
    rows = t[cta]
    base = int(rows[0, 0]) if rows[0, 0] > 0 else int(rows[0, 5])
    print(f"cta {cta}")
    for i in range(T):
        r = rows[i]
        if int(r[5]) == 0 and int(r[0]) == 0:
            break
        rel = [int(v) - base if int(v) else -1 for v in r[:7]]
        print(f"  tile {i}: mma_start {rel[0]} stage_ready {rel[1]} acc0_free {rel[2]} acc1_free {rel[3]} last_commit {rel[4]} "
              f"| epi_start {rel[5]} epi_done {rel[6]} (epi {rel[6] - rel[5]}) full_wait {int(r[7])}")

