#!/usr/bin/env python
"""Benchmark of the QLoRA linear-stack hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" = forward of all 224 QLoRA linears (32 layers x q/k/v/o/gate/up/down, NF4 base with
double-quantised absmax + LoRA r=64, dropout 0.05) followed by backward of all (dX, dA, dB into
flat buckets, all-reduced over NCCL when N > 1) on 8 x 2048 synthetic tokens per GPU
(BASELINE.json configs[2]; configs[3] is the same per GPU at N=8).  Rank 0 prints ONE JSON line.

  value     whole-job tokens/s, inputs resident in HBM, C-ABI ops called back to back
  e2e       same metric through the module surface (LoraLinear4bit.forward + autograd), with the
            step's activations copied host->device from pinned memory and the result read back
  roofline  tensor-core roofline of the dominant kernel (NF4-decode tcgen05 GEMM, fwd + dX)
  cpu_baseline / --impl reference   the CPU oracle (fp32 restatement of bitsandbytes + PEFT) on host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "qlora_linear_stack_train_tokens_per_s"
UNIT = "tokens/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"burst": float(d["bf16_tflops"]), "sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "hbm": float(d["hbm_gbs"]), "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 200 ms while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_ev = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
            }
            while not self._stop_ev.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
                self._stop_ev.wait(0.2)
        except Exception as e:  # no NVML: leave the record empty rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


CPU_SAMPLE_TOKENS = 256


def cpu_layer_sample(args, threads: int):
    """The CPU restatement of the reference algorithm (oracle/, fp32, NF4 decode in both passes as MatMul4Bit does) on ONE
    decoder layer of the workload -- all 7 projections at their real shapes, LoRA rank as configured -- over
    CPU_SAMPLE_TOKENS tokens.  The 32 layers of the stack are identical in shape, so whole-stack tokens/s =
    tokens / (layers x seconds per layer step): a bounded sample, scaled by a count, not extrapolated across shapes."""
    from importlib import import_module

    from oracle.fast import LayerSample

    stack = import_module("causal-unified-language-vision_b200.stack")
    return LayerSample(stack.SHAPE_SETS[args.shapes], args.r, CPU_SAMPLE_TOKENS, threads)


def cpu_sample_text(args):
    return (f"one decoder layer of the workload ({args.shapes}: 7 QLoRA linears, NF4 double-quant + LoRA r={args.r}) on "
            f"{CPU_SAMPLE_TOKENS} tokens, fp32 fwd+bwd (dX, dA, dB) with the NF4 decode in both passes (oracle/fast.py: the "
            f"C restatement on all host cores + torch fp32 GEMMs); stack tokens/s = {CPU_SAMPLE_TOKENS} / ({args.layers} "
            "layers x seconds per layer step)")


def workload_config(args, world):
    return {
        "workload": f"{args.layers}-layer 7B QLoRA linear stack ({args.shapes}: q/k/v/o 4096, gate/up/down 4096x14336), "
                    f"NF4 blocksize 64 + double-quant absmax, LoRA r={args.r} alpha=16 dropout={args.dropout}, "
                    f"batch {args.batch} x seq {args.seq} per GPU, fwd+bwd (dX, dA, dB)",
        "baseline_config": "configs[2]" if world == 1 else "configs[3]",
        "layers": args.layers, "tokens_per_gpu": args.batch * args.seq, "global_tokens": args.batch * args.seq * world,
        "rank_r": args.r, "dropout": args.dropout, "recompute": bool(args.recompute),
        "parallelism": f"dp{world}", "cache": "inputs_exceed_l2 (4 GB packed weights + >=134 MB activations per launch)",
    }


def run_reference(args, world, rank):
    """--impl reference: the CPU oracle on the host cores, rank 0 only."""
    if rank != 0:
        return
    from importlib import import_module

    threads = os.cpu_count() or 1
    sample = cpu_layer_sample(args, threads)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        sample.step()
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.mean(times)
    toks = CPU_SAMPLE_TOKENS / (args.layers * sec)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": toks, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": toks, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample_text(args),
                         "gflops": sample.flops / sec / 1e9},
        "e2e": {"value": toks, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--r", type=int, default=64)
    ap.add_argument("--dropout", type=float, default=0.05)
    ap.add_argument("--shapes", default="mistral_literal")
    ap.add_argument("--recompute", action="store_true", help="also re-run the forward inside backward (grad ckpt)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-opt", action="store_true", help="skip the (reported-only) fused optimizer timing")
    ap.add_argument("--e2e-timeout", type=float, default=120.0, help="seconds after which the e2e phase is abandoned")
    ap.add_argument("--global-timeout", type=float, default=420.0, help="hard limit for the whole run (any N)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, world, rank)
        return

    import torch
    import torch.distributed as dist
    from importlib import import_module

    t_start = time.perf_counter()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference "
                         "for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def dump_state(why):
        """Where is this rank?  Python stacks of every thread + the library's launch counter, on stderr (called from the
        watchdog timers just before they give up, so a hung multi-GPU run leaves evidence behind)."""
        import faulthandler
        try:
            n = sys.modules["b200qlora"].functional.launch_count() if "b200qlora" in sys.modules else -1
        except Exception:  # noqa: BLE001 -- diagnostics must not raise
            n = -1
        print(f"[bench rank {rank}] state dump ({why}): b2q launches so far {n}", file=sys.stderr, flush=True)
        print_stalls()
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        sys.stderr.flush()

    def stall_text():
        """Records of the kernels' stall guard (include/b2q.h): host memory, readable after the context is lost."""
        try:
            return sys.modules["b200qlora"]._lib.stall_report() if "b200qlora" in sys.modules else ""
        except Exception:  # noqa: BLE001 -- diagnostics must not raise
            return ""

    def print_stalls():
        t = stall_text()
        if t:
            print(f"[bench rank {rank}] {t}", file=sys.stderr, flush=True)

    # global safety net at every N: never hold a box for minutes if a rank gets stuck before `value` exists
    def global_timeout():
        print(f"[bench rank {rank}] no result after {args.global_timeout:.0f} s -- giving up", file=sys.stderr, flush=True)
        dump_state("global timeout")
        os._exit(4)
    g = threading.Timer(args.global_timeout, global_timeout)
    g.daemon = True
    g.start()
    # last resort if the main thread is stuck while holding the GIL (the timers above then never run): faulthandler's
    # watchdog is a C thread -- it dumps every Python stack and _exit(1)s without needing the interpreter
    import faulthandler
    faulthandler.dump_traceback_later(args.global_timeout + 30.0, exit=True, file=sys.stderr)

    def die(where, exc):
        """A CUDA error (e.g. the stall guard's trap) in a device phase: report and leave without touching CUDA again."""
        print(f"[bench rank {rank}] {where} failed: {type(exc).__name__}: {str(exc)[:600]}", file=sys.stderr, flush=True)
        print_stalls()
        sys.stderr.flush()
        os._exit(3)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        # fail fast instead of hanging the box if a rank gets stuck
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    import b200qlora as q

    stackmod = import_module("causal-unified-language-vision_b200.stack")
    F = q.functional
    shapes = stackmod.SHAPE_SETS[args.shapes]
    M = args.batch * args.seq
    stack = stackmod.QLoRALinearStack(args.layers, shapes, M, r=args.r, dropout=args.dropout, device=dev, seed=0)
    fpt = stack.flops_per_token()
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def mark(what):  # progress markers on stderr (one line per rank and phase): tells where a rank stopped if a run hangs
        print(f"[bench rank {rank}] {what} t={time.perf_counter() - t_start:.1f}s", file=sys.stderr, flush=True)

    mark("stack built")

    # ---- per-launch timing of the dominant kernel (events on the launching stream) -------------
    main_events, main_flops = [], [0.0]
    orig_fwd, orig_dx = F.qlora_fwd, F.qlora_bwd_dx
    timing_on = [False]

    def timed(fn, flops_of):
        def wrapper(*a):
            if not timing_on[0]:
                return fn(*a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a)
            e1.record()
            main_events.append((e0, e1))
            main_flops[0] += flops_of(*a)
            return out
        return wrapper

    def fwd_flops(x, packed, qs, us, lora_b, *rest):
        m, n, k = x.shape[0], int(qs.shape[0]), int(qs.shape[1])
        return 2.0 * m * n * k + (0.0 if us is None else 2.0 * m * us.shape[1] * n)

    def dx_flops(dy, packed, qs, du, lora_a, *rest):
        m, n, k = dy.shape[0], int(qs.shape[0]), int(qs.shape[1])
        return 2.0 * m * n * k + (0.0 if du is None else 2.0 * m * du.shape[1] * k)

    F.qlora_fwd = timed(orig_fwd, fwd_flops)
    F.qlora_bwd_dx = timed(orig_dx, dx_flops)

    # ---- value: device-resident inputs, C-ABI ops back to back -------------------------------
    try:
        for _ in range(args.warmup):
            stack.step_direct(recompute=args.recompute)
        barrier()
        mark("value warm-up done")
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = F.launch_count()
        timing_on[0] = True
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            stack.step_direct(recompute=args.recompute)
        t1.record()
        barrier()
        timing_on[0] = False
        launches = F.launch_count() - launches0
        clocks = sampler.stop()
        mark("value timed steps done")
        ms = t0.elapsed_time(t1) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
    except Exception as exc:  # noqa: BLE001 -- a CUDA error here (stall guard trap, launch failure) ends the run with a report
        die("value phase", exc)
    value = M * world / (ms / 1e3)

    kern_ms = sum(a.elapsed_time(b) for a, b in main_events)
    n_main = len(main_events)
    achieved = main_flops[0] / (kern_ms / 1e3) / 1e12 if kern_ms > 0 else 0.0
    F.qlora_fwd, F.qlora_bwd_dx = orig_fwd, orig_dx
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "main_kernel_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch_avg")

    def result_line(e2e, cpu, opt_info, note=None):
        step_tflops = fpt * M / (ms / 1e3) / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args, world),
            "per_gpu_tokens_per_s": value / world,
            "step_tflops_per_gpu": step_tflops,
            "pct_bf16_tc_peak": {"of_measured_burst": step_tflops / peaks["burst"],
                                 "of_measured_sustained": step_tflops / peaks["sustained"], "peaks": peaks["source"]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "qlora_gemm_kernel (NF4-decode tcgen05 GEMM, fwd + dX launches)",
                         "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["sustained"], "traffic": traffic, "launches_timed": n_main,
                         "peak_kind": f"bf16_tflops_sustained ({peaks['source']})",
                         "share_of_step": kern_ms / (ms * args.steps) if ms > 0 else None},
            "cpu_baseline": cpu, "optimizer_step": opt_info,
        }
        if note:
            out["note"] = note
        return json.dumps(out)

    # ---- e2e: host buffers in, result out, inside the timed region ---------------------------------
    # At every N: the step through the module surface (LoraLinear4bit.forward + autograd), activations copied from pinned
    # host memory at the start of the step, squared gradient norm read back at its end.
    e2e = None
    e2e_guard = None
    if not args.no_e2e:
        widths_in = sorted({k for _, _, k in shapes})
        widths_out = sorted({n for _, n, _ in shapes})
        base_k, base_n = widths_in[0], widths_out[0]
        host_x = stack.inputs[base_k].cpu().pin_memory()
        host_dy = stack.grads_out[base_n].cpu().pin_memory()
        host_out = torch.empty(1, dtype=torch.float32).pin_memory()
        h2d = host_x.numel() * 2 + host_dy.numel() * 2

        def widen(t, width):
            reps = (width + t.shape[1] - 1) // t.shape[1]
            return t if width == t.shape[1] else torch.cat([t] * reps, dim=1)[:, :width].contiguous()

        trace = os.environ.get("B2Q_BENCH_TRACE") == "1"   # diagnostic: synchronise + mark after every phase

        def tmark(what):
            if trace:
                torch.cuda.synchronize()
                mark("e2e: " + what)

        # Input pipeline of the e2e step: every step's activations come from pinned host memory, copied on a side stream.
        # The copy of step i+1 is issued while step i computes (what a data loader with one batch of prefetch does); every
        # step still pays for its own copy inside the timed region, it just does not serialise with the kernels.
        copy_stream = torch.cuda.Stream(device=dev)
        inflight = {}

        def issue_copy():
            with torch.cuda.stream(copy_stream):
                x = host_x.to(dev, non_blocking=True)
                dy = host_dy.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            inflight["next"] = (x, dy, ev)

        def e2e_step(modules):
            if "next" not in inflight:
                issue_copy()
            x, dy, ev = inflight.pop("next")
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            x.record_stream(cur)
            dy.record_stream(cur)
            if os.environ.get("B2Q_E2E_PREFETCH", "1") == "1":
                issue_copy()            # the next step's inputs, overlapping this step's kernels
            tmark("h2d done")
            ins = {k: widen(x, k) for k in widths_in}
            gos = {n: widen(dy, n) for n in widths_out}
            tmark("widen done")
            if modules:
                g2 = stack.step_modules(ins, gos, tmark if trace else None,
                                        interleaved=os.environ.get("B2Q_E2E_ORDER") == "per_module")
            else:
                stack.step_direct(recompute=args.recompute, inputs=ins, grads_out=gos)
                g2 = stack.grad_sqnorm()
            tmark("step done")
            host_out.copy_(g2.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(host_out[0])

        def time_e2e(modules, api, order):
            label = "modules" if modules else "c_abi"
            for i in range(2):
                e2e_step(modules)
                mark(f"e2e[{label}] warm-up step {i} done")
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                e2e_step(modules)
            e1.record()
            barrier()
            ems = e0.elapsed_time(e1) / args.steps
            mark(f"e2e[{label}] timed steps done")
            if world > 1:
                t = torch.tensor([ems], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ems = float(t.item())
            return {"value": M * world / (ems / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ems, "api": api, "order": order,
                    "input_pipeline": ("pinned host -> device on a copy stream; the copy of step i+1 overlaps the kernels of step i"
                                       if os.environ.get("B2Q_E2E_PREFETCH", "1") == "1" else
                                       "pinned host -> device on a copy stream at the start of each step")}

        mark("e2e host buffers pinned")
        API_MODULES = "LoraLinear4bit.forward + autograd (QLoRALinear) -> C ABI"
        ORDER_ALL = "forward of all modules, then backward of all"
        ORDER_PER = "per-module forward+backward, last module first (same kernels and launch count, no 50 GB of live outputs)"
        interleaved = os.environ.get("B2Q_E2E_ORDER") == "per_module"

        # The e2e phase runs under a watchdog at every N.  If it does not finish (or dies with a CUDA error -- the kernels'
        # stall guard traps after ~3 s and leaves a report), rank 0 still prints the complete device-timed line with
        # "e2e": null and a note, and the process exits non-zero within --e2e-timeout: never a box held for 30 minutes.
        def e2e_failed(why):
            mark("e2e phase failed: " + why)
            dump_state("e2e failure")
            if rank == 0:
                note = (f"e2e phase did not complete at n_gpus={world}: {why}; value / roofline are device-timed and complete. "
                        + stall_text()[:1500])
                print(result_line(None, None, None, note=note), flush=True)
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(5)

        e2e_guard = threading.Timer(args.e2e_timeout, lambda: e2e_failed(f"no result within {args.e2e_timeout:.0f} s"))
        e2e_guard.daemon = True
        e2e_guard.start()
        try:
            e2e = time_e2e(True, API_MODULES, ORDER_PER if interleaved else ORDER_ALL)
        except Exception as exc:  # noqa: BLE001
            e2e_failed(f"{type(exc).__name__}: {str(exc)[:300]}")
        e2e_guard.cancel()
        e2e_guard = None

    # ---- the step either side of the path: fused clip + AdamW on the flat buckets (reported, not part of `value`) ---
    opt_info = None
    if not args.no_opt:
        optim = import_module("causal-unified-language-vision_b200.optim")
        opt = optim.FusedLoraAdamW(stack.sync, lr=2e-5, weight_decay=0.0)
        for _ in range(2):
            opt.clip_grad_norm_(10.0)
            opt.step()
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for _ in range(5):
            opt.clip_grad_norm_(10.0)
            opt.step()
        o1.record()
        barrier()
        oms = o0.elapsed_time(o1) / 5
        nel = sum(f.numel() for f in opt.pflat)
        opt_info = {"ms": oms, "elements": nel, "GBps": (16.0 * nel) / (oms / 1e3) / 1e9,
                    "what": "global-norm clip (2 B/elem read) + AdamW (14 B/elem) on bf16 LoRA buckets, bf16 moments"}

    if e2e_guard is not None:
        e2e_guard.cancel()
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            sample = cpu_layer_sample(args, threads)
            ts = []
            for i in range(6):   # 1 warm-up + 5 timed layer steps: 10-20 s of host work
                t0c = time.perf_counter()
                sample.step()
                if i > 0:
                    ts.append(time.perf_counter() - t0c)
            sec = statistics.median(ts)
            cpu = {"value": CPU_SAMPLE_TOKENS / (args.layers * sec), "unit": UNIT, "cores": threads, "kind": "port",
                   "gflops": sample.flops / sec / 1e9, "sample": cpu_sample_text(args)}
        print(result_line(e2e, cpu, opt_info), flush=True)
    faulthandler.cancel_dump_traceback_later()
    g.cancel()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
