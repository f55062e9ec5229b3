#!/bin/bash
# Round 2, lease 15: streaming (evict-first) epilogue stores vs default write-back, same-box A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease15; mkdir -p $O
P=$PWD/causal-unified-language-vision_b200
S() { echo "$@" | tee -a $O/summary.txt; }
B2Q_LIB_PATH=$P/libb2q_stcs.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/pytest_stcs.log 2>&1; S "pytest stcs rc=$? $(tail -1 $O/pytest_stcs.log)"
for rep in 1 2 3; do for v in default stcs; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_${v}_$rep.out 2> $O/ab_${v}_$rep.err
  S "A/B $v $rep rc=$? $(grep -o '"value": [0-9.]*' $O/ab_${v}_$rep.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/ab_${v}_$rep.out | head -1)"
done; done
for v in default stcs; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 300 python tests/gpu_step_breakdown.py > $O/breakdown_$v.txt 2>&1; S "$v $(grep -E 'step|lora_down|qlora_fwd|lora_bwd_du|qlora_bwd_dx|lora_grads' $O/breakdown_$v.txt | tr -s ' ' | tr '\n' ' ')"
done
