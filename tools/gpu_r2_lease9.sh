#!/bin/bash
# Round 2, lease 9: backward call order A/B (lora_grads before / after qlora_bwd_dx), soak, repeated full bench runs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease9; mkdir -p $O
S() { echo "$@" | tee -a $O/summary.txt; }
timeout 500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; S "pytest rc=$? $(tail -1 $O/pytest.log)"
for rep in 1 2 3; do for v in 0 1; do
  B2Q_GRADS_BEFORE_DX=$v timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_gfirst${v}_$rep.out 2> $O/ab_gfirst${v}_$rep.err
  S "A/B grads_before_dx=$v $rep rc=$? $(grep -o '"value": [0-9.]*' $O/ab_gfirst${v}_$rep.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/ab_gfirst${v}_$rep.out | head -1)"
done; done
for v in 0 1; do B2Q_GRADS_BEFORE_DX=$v timeout 300 python tests/gpu_step_breakdown.py > $O/breakdown_gfirst$v.txt 2>&1; S "grads_before_dx=$v $(grep -E 'step|lora_grads|qlora_bwd_dx' $O/breakdown_gfirst$v.txt | tr '\n' ' ')"; done
timeout 400 python tools/stall_hunt.py --iters 60 --watchdog 380 > $O/soak.out 2> $O/soak.err; S "soak (60 x [2 C-ABI steps + fresh-allocation module-surface step]) rc=$? $(tail -1 $O/soak.err | cut -c1-140)"
for i in 1 2 3 4; do
  timeout 240 python bench.py --steps 20 --warmup 5 > $O/bench_$i.out 2> $O/bench_$i.err; S "bench $i rc=$? $(grep -o '"value": [0-9.]*' $O/bench_$i.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench_$i.out | head -1)"
done
