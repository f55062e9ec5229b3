"""Bring-up probe for the CUDA kernels (not a pytest file): each step runs in its own subprocess
under a timeout so a hung kernel cannot stall the GPU box, and appends to gpurun_out/probe.log.

    python tests/gpu_probe.py            # run every step
    python tests/gpu_probe.py STEP ...   # run the named steps in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STEPS = ["decode", "quantize", "gemm_nk", "gemm_kn", "lora_down", "lora_du", "lora_grads",
         "fwd0", "fwd1", "fwd2", "fwd3", "dx0", "dx1", "dx2", "dx3", "fwd_big", "dx_big"]


def _state_to_gpu(state, dev):
    import torch
    import b200qlora as q

    code = torch.from_numpy(state["code"]).to(dev)
    packed = torch.from_numpy(state["packed"]).to(dev).reshape(-1, 1)
    if state["nested"]:
        s2 = q.QuantState(torch.from_numpy(state["absmax2"]).to(dev), code=torch.from_numpy(state["code256"]).to(dev),
                          blocksize=256, quant_type="dynamic8", dtype=torch.float32)
        qs = q.QuantState(torch.from_numpy(state["absmax_q"]).to(dev), state["shape"], code, 64, "nf4",
                          torch.bfloat16, offset=torch.tensor(float(state["offset"]), device=dev), state2=s2)
    else:
        qs = q.QuantState(torch.from_numpy(state["absmax"]).to(dev), state["shape"], code, 64, "nf4", torch.bfloat16)
    return packed, qs


def run_step(name):
    import numpy as np
    import torch
    import b200qlora as q
    import oracle
    from oracle import nf4
    from oracle.qlora import make_case, rel_err

    F = q.functional
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    ok = True

    def report(tag, err, tol):
        nonlocal ok
        good = err <= tol
        ok &= good
        print(f"  {tag}: err={err:.3e} tol={tol:.1e} {'OK' if good else 'FAIL'}", flush=True)

    if name == "decode":
        rng = np.random.default_rng(0)
        W = rng.normal(0, 0.02, (512, 1024)).astype(np.float32)
        W[3, :64] = 0.0
        for dq in (False, True):
            st = nf4.quantize_nf4(W, 64, dq)
            ref = torch.from_numpy(nf4.dequantize_nf4(st).astype(np.int16))
            packed, qs = _state_to_gpu(st, dev)
            for algo in (0, 1):
                out = F.dequantize_4bit(packed, qs, algo=algo)
                same = torch.equal(out.view(torch.int16).cpu(), ref)
                ok &= same
                print(f"  decode nested={dq} algo={algo}: bit-exact={same}", flush=True)
    elif name == "quantize":
        rng = np.random.default_rng(1)
        W = rng.normal(0, 0.02, (256, 512)).astype(np.float32)
        W[5, 64:128] = 0.0
        for dq in (False, True):
            st = nf4.quantize_nf4(W, 64, dq)
            packed, qs = F.quantize_4bit(torch.from_numpy(W).to(dev), compress_statistics=dq)
            same = np.array_equal(packed.cpu().numpy().reshape(-1), st["packed"])
            if dq:
                same &= np.array_equal(qs.absmax.cpu().numpy(), st["absmax_q"])
                same &= np.array_equal(qs.state2.absmax.cpu().numpy(), st["absmax2"])
                same &= float(qs.offset.item()) == float(st["offset"])
            else:
                same &= np.array_equal(qs.absmax.cpu().numpy(), st["absmax"])
            ok &= bool(same)
            print(f"  quantize nested={dq}: identical={bool(same)}", flush=True)
    elif name in ("gemm_nk", "gemm_kn"):
        for (M, N, K) in ((256, 256, 128), (300, 384, 512), (1024, 1024, 4096)):
            a = torch.randn(M, K, device=dev).bfloat16()
            if name == "gemm_nk":
                b = torch.randn(N, K, device=dev).bfloat16()
                ref = a.float() @ b.float().t()
                d = F.gemm_bf16(a, b, False)
            else:
                b = torch.randn(K, N, device=dev).bfloat16()
                ref = a.float() @ b.float()
                d = F.gemm_bf16(a, b, True)
            torch.cuda.synchronize()
            report(f"{name} {M}x{N}x{K}", rel_err(d.cpu(), ref.cpu()), 1e-2)
    elif name == "lora_down":
        for (M, K, r) in ((704, 4096, 64), (512, 1024, 128)):
            x = torch.randn(M, K, device=dev).bfloat16()
            A = (torch.rand(r, K, device=dev) * 2 - 1).mul(K ** -0.5).bfloat16()
            u, us = F.lora_down(x, A, 0.25)
            ref = x.float() @ A.float().t()
            report(f"lora_down u {M}x{K}x{r}", rel_err(u.cpu(), ref.cpu()), 1e-2)
            report(f"lora_down us {M}x{K}x{r}", rel_err(us.cpu(), 0.25 * ref.cpu()), 1e-2)
    elif name == "lora_du":
        for (M, N, r) in ((704, 4096, 64), (512, 1024, 128)):
            dy = torch.randn(M, N, device=dev).bfloat16()
            B = (torch.randn(N, r, device=dev) * 0.02).bfloat16()
            du = F.lora_bwd_du(dy, B, 0.25)
            ref = 0.25 * (dy.float() @ B.float())
            report(f"lora_du {M}x{N}x{r}", rel_err(du.cpu(), ref.cpu()), 1e-2)
    elif name == "lora_grads":
        for (M, N, K, r) in ((704, 1024, 512, 64), (2048, 4096, 4096, 64), (1000, 512, 1024, 128)):
            dy = torch.randn(M, N, device=dev).bfloat16()
            x = torch.randn(M, K, device=dev).bfloat16()
            u = torch.randn(M, r, device=dev).bfloat16()
            du = torch.randn(M, r, device=dev).bfloat16()
            dA = torch.zeros(r, K, device=dev, dtype=torch.bfloat16)
            dB = torch.zeros(N, r, device=dev, dtype=torch.bfloat16)
            F.lora_grads(dy, x, u, du, 0.25, dA, dB)
            refA = du.float().t() @ x.float()
            refB = 0.25 * (dy.float().t() @ u.float())
            report(f"dA {M}x{N}x{K}x{r}", rel_err(dA.cpu(), refA.cpu()), 1e-2)
            report(f"dB {M}x{N}x{K}x{r}", rel_err(dB.cpu(), refB.cpu()), 1e-2)
    elif name.startswith("fwd") or name.startswith("dx"):
        big = name.endswith("_big")
        is_fwd = name.startswith("fwd")
        variant = 3 if big else int(name[-1])
        shapes = ((2048, 4096, 4096, 64),) if big else ((512, 512, 512, 64), (704, 1024, 2048, 64), (200, 256, 256, 128))
        for (M, N, K, r) in shapes:
            for dq in (False, True):
                for lora in (False, True):
                    case = make_case(M, N, K, r, seed=M + N, double_quant=dq)
                    st = case["state"]
                    packed, qs = _state_to_gpu(st, dev)
                    W = torch.from_numpy(nf4.dequantize_nf4(st, as_bits=False).copy())
                    x, dy, A, B = (case[k].to(dev) for k in ("x", "dy", "A", "B"))
                    if is_fwd:
                        F.set_variant(variant, -1)
                        if lora:
                            u, us = F.lora_down(x, A, 0.25)
                            y = F.qlora_fwd(x, packed, qs, us, B)
                            ref = x.float().cpu() @ W.t() + us.float().cpu() @ B.float().cpu().t()
                        else:
                            y = F.qlora_fwd(x, packed, qs, None, None)
                            ref = x.float().cpu() @ W.t()
                        torch.cuda.synchronize()
                        report(f"{name} {M}x{N}x{K} r={r} nested={dq} lora={lora}", rel_err(y.cpu(), ref), 1e-2)
                    else:
                        F.set_variant(-1, variant)
                        if lora:
                            du = F.lora_bwd_du(dy, B, 0.25)
                            dx = F.qlora_bwd_dx(dy, packed, qs, du, A)
                            ref = dy.float().cpu() @ W + du.float().cpu() @ A.float().cpu()
                        else:
                            dx = F.qlora_bwd_dx(dy, packed, qs, None, None)
                            ref = dy.float().cpu() @ W
                        torch.cuda.synchronize()
                        report(f"{name} {M}x{N}x{K} r={r} nested={dq} lora={lora}", rel_err(dx.cpu(), ref), 1e-2)
    else:
        raise SystemExit(f"unknown step {name}")
    torch.cuda.synchronize()
    print(f"STEP {name}: {'PASS' if ok else 'FAIL'}", flush=True)
    return ok


def main():
    if len(sys.argv) > 1 and sys.argv[1] != "--all":
        good = True
        for s in sys.argv[1:]:
            good &= run_step(s)
        sys.exit(0 if good else 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "probe.log"), "a")
    summary = []
    for s in STEPS:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), s], capture_output=True, text=True,
                               timeout=240, cwd=ROOT)
            out, status = r.stdout + r.stderr[-3000:], ("PASS" if r.returncode == 0 else f"FAIL rc={r.returncode}")
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
            status = "TIMEOUT"
        line = f"== {s}: {status} ({time.time() - t0:.1f}s)"
        print(line, flush=True)
        log.write(line + "\n" + out + "\n")
        log.flush()
        summary.append(line)
    print("\n".join(summary))


if __name__ == "__main__":
    main()
