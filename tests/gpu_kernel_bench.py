"""Per-kernel timing on one B200 (CUDA events, warm-up, L2-exceeding working sets).

    python tests/gpu_kernel_bench.py [M] [--variants 0,1,2,3] [--shapes 4096x4096,...]
Prints one line per (kernel, shape, variant): ms and TFLOP/s (algorithmic) or GB/s.
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


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("M", nargs="?", type=int, default=16384)
    ap.add_argument("--variants", default="0,1,2,3")
    ap.add_argument("--shapes", default="4096x4096,14336x4096,4096x14336")
    ap.add_argument("--r", type=int, default=64)
    ap.add_argument("--skinny-only", action="store_true", help="only the HBM-bound LoRA kernels (with and without dropout)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernel_bench.jsonl"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    M, r = args.M, args.r
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "a")

    def emit(**kw):
        print(json.dumps(kw), flush=True)
        out.write(json.dumps(kw) + "\n")
        out.flush()

    for shp in args.shapes.split(","):
        N, K = (int(v) for v in shp.split("x"))
        W = torch.randn(N, K, device=dev) * 0.02
        packed, qs = F.quantize_4bit(W, compress_statistics=True)
        Wb = F.dequantize_4bit(packed, qs)
        del W
        x = torch.randn(M, K, device=dev).bfloat16()
        dy = (torch.randn(M, N, device=dev) / N ** 0.5).bfloat16()
        A = ((torch.rand(r, K, device=dev) * 2 - 1) / K ** 0.5).bfloat16()
        B = (torch.randn(N, r, device=dev) * 0.02).bfloat16()
        gemm_flops = 2.0 * M * N * K
        # cuBLAS reference points (materialised bf16 weight)
        med, best = timeit(lambda: torch.matmul(x, Wb.t()))
        emit(kernel="cublas_fwd", N=N, K=K, M=M, ms=med, best_ms=best, tflops=gemm_flops / med / 1e9)
        med, best = timeit(lambda: torch.matmul(dy, Wb))
        emit(kernel="cublas_dx", N=N, K=K, M=M, ms=med, best_ms=best, tflops=gemm_flops / med / 1e9)
        med, best = timeit(lambda: F.dequantize_4bit(packed, qs))
        emit(kernel="nf4_decode", N=N, K=K, ms=med, gbs=(N * K * 2.5) / med / 1e6)
        med, best = timeit(lambda: F.lora_down(x, A, 0.25))
        emit(kernel="lora_down", N=N, K=K, M=M, ms=med, gbs=(2.0 * M * K) / med / 1e6)
        u, us = F.lora_down(x, A, 0.25)
        med, best = timeit(lambda: F.lora_bwd_du(dy, B, 0.25))
        emit(kernel="lora_bwd_du", N=N, K=K, M=M, ms=med, gbs=(2.0 * M * N) / med / 1e6)
        du = F.lora_bwd_du(dy, B, 0.25)
        dA = torch.zeros(r, K, device=dev, dtype=torch.bfloat16)
        dB = torch.zeros(N, r, device=dev, dtype=torch.bfloat16)
        med, best = timeit(lambda: F.lora_grads(dy, x, u, du, 0.25, dA, dB))
        emit(kernel="lora_grads", N=N, K=K, M=M, ms=med, gbs=(2.0 * M * (N + K)) / med / 1e6)
        med, best = timeit(lambda: F.lora_down(x, A, 0.25, 99, 0.05))
        emit(kernel="lora_down_drop", N=N, K=K, M=M, ms=med, gbs=(2.0 * M * K) / med / 1e6)
        med, best = timeit(lambda: F.lora_grads(dy, x, u, du, 0.25, dA, dB, seed=99, p=0.05))
        emit(kernel="lora_grads_drop", N=N, K=K, M=M, ms=med, gbs=(2.0 * M * (N + K)) / med / 1e6)
        if args.skinny_only:
            del x, dy, Wb, packed
            torch.cuda.empty_cache()
            continue
        for v in (int(s) for s in args.variants.split(",")):
            F.set_variant(v, v)
            for lora in (False, True):
                fl = gemm_flops + (2.0 * M * r * N if lora else 0)
                med, best = timeit(lambda: F.qlora_fwd(x, packed, qs, us if lora else None, B if lora else None))
                emit(kernel="qlora_fwd", variant=v, lora=lora, N=N, K=K, M=M, ms=med, best_ms=best,
                     tflops=fl / med / 1e9)
                fl = gemm_flops + (2.0 * M * r * K if lora else 0)
                med, best = timeit(lambda: F.qlora_bwd_dx(dy, packed, qs, du if lora else None, A if lora else None))
                emit(kernel="qlora_bwd_dx", variant=v, lora=lora, N=N, K=K, M=M, ms=med, best_ms=best,
                     tflops=fl / med / 1e9)
                if lora:  # LoRA dropout p = 0.05 (the training configuration)
                    med, best = timeit(lambda: F.qlora_bwd_dx(dy, packed, qs, du, A, 1234, 0.05))
                    emit(kernel="qlora_bwd_dx_drop", variant=v, lora=lora, N=N, K=K, M=M, ms=med, best_ms=best,
                         tflops=fl / med / 1e9)
        del x, dy, Wb, packed
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
