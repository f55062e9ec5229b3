#!/usr/bin/env python
"""Which k-block does a corrupted tile come from?  Activation rows are non-zero in ONE 64-wide k-block only (row m uses
k-block m % KB), so every output row tests exactly one k-block of the decode pipeline: a stale / early-read operand stage
shows up as (tile, W-row band, k-block).  Development tool (GPU)."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import b200qlora as q  # noqa: E402

F = q.functional
dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = 4096
tot_bad = 0
for N, K in ((4096, 4096), (14336, 4096)):
    g = torch.Generator(device=dev).manual_seed(N + K)
    packed, qs = F.quantize_4bit(torch.empty(N, K, device=dev).normal_(0, 0.02, generator=g), compress_statistics=True)
    W = F.dequantize_4bit(packed, qs)
    for direction in ("fwd", "dx"):
        C = K if direction == "fwd" else N          # contraction length
        KB = C // 64
        a = torch.zeros(M, C, device=dev)
        rows = torch.arange(M, device=dev)
        blk = rows % KB
        vals = torch.empty(M, 64, device=dev).normal_(generator=g)
        idx = (blk * 64).unsqueeze(1) + torch.arange(64, device=dev).unsqueeze(0)
        a.scatter_(1, idx, vals)
        a = a.bfloat16()
        ref = (a @ (W.t() if direction == "fwd" else W)).float()
        sc = float(ref.abs().max())
        events = collections.Counter()
        for it in range(reps):
            got = (F.qlora_fwd(a, packed, qs, None, None) if direction == "fwd"
                   else F.qlora_bwd_dx(a, packed, qs, None, None)).float()
            bad = (got - ref).abs() > 0.02 * sc
            nb = int(bad.sum())
            if nb == 0:
                continue
            tot_bad += 1
            br, bc = bad.nonzero(as_tuple=True)
            kbs = (br % KB).tolist()
            tiles = (br // 512).tolist()
            bands = (bc // 32).tolist()
            ev = collections.Counter(zip(tiles, bands, kbs))
            for (t, b, k), c in sorted(ev.items()):
                if c >= 4:
                    events[(k % 4, k)] += 1
                    print(f"{direction} N={N} K={K} it={it}: m-tile {t} cols {b * 32}..{b * 32 + 31} (col%256={b * 32 % 256}) "
                          f"k-block {k}/{KB} (stage {k % 4}) {c} elements", flush=True)
        print(f"== {direction} N={N} K={K}: bad launches so far {tot_bad}; k-block histogram {sorted(events.items())[:40]}", flush=True)
print("bad launches:", tot_bad)
