"""GEMV (single-token decode) timing on one B200: GB/s of packed-weight bytes vs the measured HBM peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import b200qlora as q  # noqa: E402

F = q.functional
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for N, K in ((4096, 4096), (14336, 4096), (4096, 14336)):
    packed, qs = F.quantize_4bit(torch.randn(N, K, device=dev) * 0.02, compress_statistics=True)
    for M in (1, 4, 8):
        x = torch.randn(M, K, device=dev).bfloat16()
        ts = []
        for i in range(8):
            flush.zero_()  # evict the weight from L2 between iterations
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            F.gemv_4bit(x, packed, qs)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        by = N * K / 2 + N * K / 64
        print(json.dumps({"kernel": "gemv_4bit", "M": M, "N": N, "K": K, "us": ms * 1e3, "GBps": by / ms / 1e6}), flush=True)
