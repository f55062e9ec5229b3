#!/bin/bash
# Round 2, lease 7: masked dX GEMM -- epilogue sets (2 vs 3) x mask hash before / after the TMEM load wait, same-box A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease7; mkdir -p $O
P=$PWD/causal-unified-language-vision_b200
S() { echo "$@" | tee -a $O/summary.txt; }
for v in default e2 e2h e3h; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dropout or full_size or c1 or repeated" > $O/pytest_$v.log 2>&1; S "pytest $v rc=$? $(tail -1 $O/pytest_$v.log)"
done
for rep in 1 2; do for v in e2 default e2h e3h; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_${v}_$rep.out 2> $O/ab_${v}_$rep.err
  S "A/B $v $rep rc=$? $(grep -o '"value": [0-9.]*' $O/ab_${v}_$rep.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/ab_${v}_$rep.out | head -1)"
done; done
unset B2Q_LIB_PATH
timeout 300 python tests/gpu_step_breakdown.py > $O/breakdown_default.txt 2>&1; cat $O/breakdown_default.txt | tee -a $O/summary.txt
