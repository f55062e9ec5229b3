#!/usr/bin/env python
"""Repeat the full-size dX parity case (14336x4096 and 4096x14336, M = 4096, dropout) and, when an element is off,
say WHERE (row / column blocks), so that a racy tile shows up as a pattern.  Development tool (GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import b200qlora as q  # noqa: E402

F = q.functional
dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
bad_total = 0
for N, K in ((14336, 4096), (4096, 14336), (4096, 4096)):
    M, r, s, p, seed = 4096, 64, 0.25, 0.05, 4242
    g = torch.Generator(device=dev).manual_seed(N)
    packed, qs = F.quantize_4bit(torch.empty(N, K, device=dev).normal_(0, 0.02, generator=g), compress_statistics=True)
    x = torch.empty(M, K, device=dev).normal_(generator=g).bfloat16()
    dy = (torch.empty(M, N, device=dev).normal_(generator=g) / N ** 0.5).bfloat16()
    A = ((torch.rand(r, K, device=dev, generator=g) * 2 - 1) / K ** 0.5).bfloat16()
    B = torch.empty(N, r, device=dev).normal_(0, 0.02, generator=g).bfloat16()
    W = F.dequantize_4bit(packed, qs)
    mask = F.dropout_mask((M, K), seed, p, dev).to(torch.bfloat16)
    du_ref = ((dy.float() * s).bfloat16() @ B)
    dx_ref = (dy @ W).float() + (du_ref @ A).float() * mask.float() / (1 - p)
    y_ref = (x @ W.t()).float()
    scale = float(dx_ref.abs().max())
    for it in range(reps):
        for variant in ("dx_drop", "dx_nodrop", "dx_base", "fwd_base"):
            if variant == "fwd_base":
                got = F.qlora_fwd(x, packed, qs, None, None).float()
                ref = y_ref
            else:
                du = F.lora_bwd_du(dy, B, s)
                if variant == "dx_drop":
                    got, ref = F.qlora_bwd_dx(dy, packed, qs, du, A, seed, p).float(), dx_ref
                elif variant == "dx_nodrop":
                    got = F.qlora_bwd_dx(dy, packed, qs, du, A, seed, 0.0).float()
                    ref = (dy @ W).float() + (du_ref @ A).float()
                else:
                    got, ref = F.qlora_bwd_dx(dy, packed, qs, None, None).float(), (dy @ W).float()
            err = (got - ref).abs()
            sc = float(ref.abs().max())
            rel = float(err.max()) / sc
            flag = "" if rel <= 2e-2 else "  <-- BAD"
            line = f"N={N} K={K} it={it} {variant:10s} max rel err {rel:.4f}{flag}"
            if rel > 2e-2:
                bad_total += 1
                bad = err > 0.01 * sc
                rows = bad.any(dim=1).nonzero().flatten()
                cols = bad.any(dim=0).nonzero().flatten()
                line += (f" | {int(bad.sum())} elements off; rows {int(rows.min())}..{int(rows.max())} ({rows.numel()} rows, "
                         f"128-blocks {sorted(set((rows // 128).tolist()))[:12]}) cols {int(cols.min())}..{int(cols.max())} "
                         f"({cols.numel()} cols, 64-blocks {sorted(set((cols // 64).tolist()))[:12]})")
            print(line, flush=True)
print("bad cases:", bad_total)
