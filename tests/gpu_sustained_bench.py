"""Sustained-regime timing of the two main kernels on one B200: each configuration runs back to back for
~0.4 s after a 0.2 s warm loop (the chip is power-capped within milliseconds, so 10-iteration bursts say
little about in-step behaviour); configurations are visited round-robin twice.

    python tests/gpu_sustained_bench.py --variants 3,4,5 [--shapes 4096x4096,...] [--drop 0.05]
Development tool; the judged numbers come from bench.py.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import b200qlora as q  # noqa: E402

F = q.functional


def run_for(fn, seconds):
    """Run fn back to back for about `seconds`; returns ms per call (CUDA events)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 5
    n = max(5, int(seconds * 1e3 / per))
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=16384)
    ap.add_argument("--variants", default="3,4,5")
    ap.add_argument("--shapes", default="4096x4096,14336x4096,4096x14336")
    ap.add_argument("--r", type=int, default=64)
    ap.add_argument("--drop", type=float, default=0.05)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=0.4)
    ap.add_argument("--cublas", action="store_true")
    ap.add_argument("--pf", default="0", help="comma list of A-operand L2 prefetch distances to sweep")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sustained_bench.jsonl"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    M, r = args.M, args.r
    out = open(args.out, "a")
    variants = [int(v) for v in args.variants.split(",")]
    for shp in args.shapes.split(","):
        N, K = (int(v) for v in shp.split("x"))
        packed, qs = F.quantize_4bit(torch.randn(N, K, device=dev) * 0.02, compress_statistics=True)
        x = torch.randn(M, K, device=dev).bfloat16()
        dy = (torch.randn(M, N, device=dev) / N ** 0.5).bfloat16()
        A = ((torch.rand(r, K, device=dev) * 2 - 1) / K ** 0.5).bfloat16()
        B = (torch.randn(N, r, device=dev) * 0.02).bfloat16()
        u, us = F.lora_down(x, A, 0.25)
        du = F.lora_bwd_du(dy, B, 0.25)
        cfgs = []
        if args.cublas:
            Wb = F.dequantize_4bit(packed, qs)
            cfgs.append(("cublas_fwd", -1, lambda: torch.matmul(x, Wb.t()), 2.0 * M * N * K))
        lib = q._lib.load()
        for pf in (int(s) for s in args.pf.split(",")):
            for v in variants:
                tag = f"{v}" if args.pf == "0" else f"{v}/pf{pf}"
                cfgs.append(("fwd", tag, lambda v=v, pf=pf: (lib.b2q_debug_set_prefetch(pf), F.set_variant(v, v),
                                                           F.qlora_fwd(x, packed, qs, us, B)), 2.0 * M * N * (K + r)))
                cfgs.append(("dx", tag, lambda v=v, pf=pf: (lib.b2q_debug_set_prefetch(pf), F.set_variant(v, v),
                                                          F.qlora_bwd_dx(dy, packed, qs, du, A)), 2.0 * M * K * (N + r)))
                if args.drop > 0:
                    cfgs.append(("dx_drop", tag, lambda v=v, pf=pf: (lib.b2q_debug_set_prefetch(pf), F.set_variant(v, v),
                                                                   F.qlora_bwd_dx(dy, packed, qs, du, A, 77, args.drop)),
                                 2.0 * M * K * (N + r)))
        acc = {}
        for rnd in range(args.rounds):
            for name, v, fn, fl in cfgs:
                run_for(fn, 0.2)
                ms = run_for(fn, args.seconds)
                acc.setdefault((name, v), []).append(fl / ms / 1e9)
        for (name, v), vals in acc.items():
            rec = {"kernel": name, "variant": v, "N": N, "K": K, "M": M, "tflops_sustained": sum(vals) / len(vals),
                   "rounds": [round(t, 1) for t in vals]}
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
        F.set_variant(-1, -1)
        del x, dy, packed
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
