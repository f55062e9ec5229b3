#!/bin/bash
# Round 2, lease 16: streaming output stores (forward always, dX without dropout) vs none, same-box A/B; full GPU suite.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease16; mkdir -p $O
P=$PWD/causal-unified-language-vision_b200
S() { echo "$@" | tee -a $O/summary.txt; }
timeout 500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; S "pytest rc=$? $(tail -1 $O/pytest.log)"
timeout 200 python tools/dx_check.py 2 > $O/dx.log 2>&1; S "dx_check rc=$? $(tail -1 $O/dx.log)"
for rep in 1 2 3; do for v in nostream default; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_${v}_$rep.out 2> $O/ab_${v}_$rep.err
  S "A/B $v $rep rc=$? $(grep -o '"value": [0-9.]*' $O/ab_${v}_$rep.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/ab_${v}_$rep.out | head -1)"
done; done
unset B2Q_LIB_PATH
timeout 240 python bench.py --steps 20 --warmup 5 > $O/bench_full.out 2> $O/bench_full.err; S "bench full rc=$? $(grep -o '"value": [0-9.]*' $O/bench_full.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench_full.out | head -1)"
