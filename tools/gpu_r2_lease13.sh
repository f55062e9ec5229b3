#!/bin/bash
# Round 2, lease 13: new GPU test (non-LoRA trainables), e2e with / without input prefetch, other BASELINE configs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease13; mkdir -p $O
S() { echo "$@" | tee -a $O/summary.txt; }
timeout 500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; S "pytest rc=$? $(tail -1 $O/pytest.log)"; grep -E "FAILED|Error" $O/pytest.log | head -5
for v in 0 1; do
  B2Q_E2E_PREFETCH=$v timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu --no-opt > $O/bench_prefetch$v.out 2> $O/bench_prefetch$v.err
  S "bench prefetch=$v rc=$? $(grep -o '"value": [0-9.]*' $O/bench_prefetch$v.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench_prefetch$v.out | head -1)"
done
timeout 240 python bench.py --steps 20 --warmup 5 > $O/bench_full.out 2> $O/bench_full.err; S "bench full rc=$? $(grep -o '"value": [0-9.]*' $O/bench_full.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench_full.out | head -1)"
# BASELINE configs[1]: one Mistral-7B-shaped decoder layer, 1 x seq 2048;  configs[4]: 4 x 576 image tokens, seq 4096, r = 128
timeout 240 python bench.py --layers 1 --batch 1 --seq 2048 --steps 50 --warmup 10 --no-cpu --no-opt > $O/bench_c2.out 2> $O/bench_c2.err; S "configs[1] one layer seq 2048 rc=$? $(grep -o '"value": [0-9.]*' $O/bench_c2.out | head -1) $(grep -o '"step_tflops_per_gpu": [0-9.]*' $O/bench_c2.out | head -1)"
timeout 300 python bench.py --batch 4 --seq 4096 --r 128 --steps 10 --warmup 3 --no-cpu --no-opt > $O/bench_c5.out 2> $O/bench_c5.err; S "configs[4] seq 4096 r=128 rc=$? $(grep -o '"value": [0-9.]*' $O/bench_c5.out | head -1) $(grep -o '"step_tflops_per_gpu": [0-9.]*' $O/bench_c5.out | head -1)"
timeout 300 python bench.py --recompute --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/bench_recompute.out 2> $O/bench_recompute.err; S "with checkpoint recompute rc=$? $(grep -o '"value": [0-9.]*' $O/bench_recompute.out | head -1)"
