#!/bin/bash
# Round 2, lease 5: new dropout-backward dataflow (keep-scaled du, tail k-block + sparse correction), lighter load fence A/B,
# in-step breakdown, real-size HF decoder, GEMV configurations.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease5; mkdir -p $O
P=$PWD/causal-unified-language-vision_b200
S() { echo "$@" | tee -a $O/summary.txt; }
timeout 500 python -m pytest tests -m gpu -q > $O/pytest_1.log 2>&1; S "pytest rc=$? $(tail -1 $O/pytest_1.log)"
for v in default fencelight; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 200 python tools/kb_probe.py 8 > $O/kb_$v.log 2>&1; S "kb_probe $v rc=$? $(tail -1 $O/kb_$v.log)"
  timeout 200 python tools/dx_check.py 3 > $O/dx_$v.log 2>&1; S "dx_check $v rc=$? $(tail -1 $O/dx_$v.log)"
done
for rep in 1 2; do for v in default fencelight; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_${v}_$rep.out 2> $O/ab_${v}_$rep.err
  S "A/B $v $rep rc=$? $(grep -o '"value": [0-9.]*' $O/ab_${v}_$rep.out | head -1) $(grep -o '"achieved": [0-9.]*' $O/ab_${v}_$rep.out | head -1)"
done; done
unset B2Q_LIB_PATH
timeout 300 python tests/gpu_step_breakdown.py > $O/breakdown.txt 2>&1; S "breakdown rc=$?"; cat $O/breakdown.txt | tee -a $O/summary.txt
timeout 600 python tools/hf_real_size.py > $O/hf.json 2> $O/hf.err; S "hf rc=$? $(cut -c1-400 $O/hf.json)"; tail -4 $O/hf.err
for cfg in 1 3; do B2Q_GEMV_CFG=$cfg timeout 200 python tests/gpu_gemv_bench.py > $O/gemv_cfg$cfg.jsonl 2> $O/gemv_cfg$cfg.err; S "gemv cfg=$cfg rc=$?"; tail -4 $O/gemv_cfg$cfg.jsonl | cut -c1-200 | tee -a $O/summary.txt; done
B2Q_GEMV_CFG=3 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k gemv > $O/gemv3_parity.log 2>&1; S "gemv cfg 3 parity rc=$? $(tail -1 $O/gemv3_parity.log)"
timeout 240 python bench.py --steps 20 --warmup 5 > $O/bench_full.out 2> $O/bench_full.err; S "bench full rc=$? $(grep -o '"value": [0-9.]*' $O/bench_full.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench_full.out | head -1)"
