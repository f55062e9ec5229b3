#!/bin/bash
# First 1-GPU call of the next round: confirm the tree is green, then run the experiments that were prepared without
# hardware (DESIGN.md "Next") -- each behind its own timeout so that a hang costs minutes, not the box.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
set -u
mkdir -p gpurun_out
step() { echo "== $1"; shift; timeout "$@"; echo "   rc=$?"; }
step "gpu tests (default configuration)" 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu.log
step "smoke" 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
# experiments: parity first, timing only if parity holds
B2Q_EXPERIMENTAL=1 step "variant 6 parity" 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variant6 or linear_fwd_bwd" > gpurun_out/r2_v6_parity.log 2>&1; tail -2 gpurun_out/r2_v6_parity.log
B2Q_EXPERIMENTAL=1 step "packed-mask parity" 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k mask_bits > gpurun_out/r2_maskbits_parity.log 2>&1; tail -2 gpurun_out/r2_maskbits_parity.log
B2Q_GEMV_CFG=3 step "tensor-core GEMV parity" 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k gemv > gpurun_out/r2_gemv3_parity.log 2>&1; tail -2 gpurun_out/r2_gemv3_parity.log
B2Q_DX_MASK_FIRST=1 step "mask-first dX parity" 300 python -m pytest tests -m gpu -x -q -k "dropout or hf or mlp" > gpurun_out/r2_maskfirst_parity.log 2>&1; tail -2 gpurun_out/r2_maskfirst_parity.log
B2Q_EXPERIMENTAL=1 step "plain-C host example" 200 python -m pytest tests/test_gpu_c_host.py -m gpu -x -q > gpurun_out/r2_c_host.log 2>&1; tail -2 gpurun_out/r2_c_host.log
for cfg in 1 3; do B2Q_GEMV_CFG=$cfg step "GEMV timing cfg=$cfg" 200 python tests/gpu_gemv_bench.py > gpurun_out/r2_gemv_cfg$cfg.jsonl 2> gpurun_out/r2_gemv_cfg$cfg.err; tail -3 gpurun_out/r2_gemv_cfg$cfg.jsonl; done
rm -f gpurun_out/sustained_bench.jsonl
step "sustained kernels, variants 5 vs 6" 600 python tests/gpu_sustained_bench.py --variants 5,6 > gpurun_out/r2_sb_v56.log 2>&1; cp gpurun_out/sustained_bench.jsonl gpurun_out/r2_sb_v56.jsonl 2>/dev/null
for v in 5 6; do step "phase trace fwd variant $v" 200 python tests/gpu_trace.py fwd $v > gpurun_out/r2_trace_fwd_v$v.log 2>&1; done
step "bench (default)" 600 python bench.py --no-cpu > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
B2Q_FWD_VARIANT=6 B2Q_DX_VARIANT=6 step "bench (variant 6)" 600 python bench.py --no-cpu --no-opt > gpurun_out/r2_bench_v6.json 2> gpurun_out/r2_bench_v6.err
B2Q_DX_MASK_FIRST=1 step "bench (mask-first dX)" 600 python bench.py --no-cpu --no-opt > gpurun_out/r2_bench_maskfirst.json 2> gpurun_out/r2_bench_maskfirst.err
B2Q_MASK_BITS=1 step "bench (packed mask)" 600 python bench.py --no-cpu --no-opt > gpurun_out/r2_bench_maskbits.json 2> gpurun_out/r2_bench_maskbits.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, round(d["value"]), "tok/s", round(d["ms_per_step"], 1), "ms", d["clocks"].get("sm_mhz"))
PY
