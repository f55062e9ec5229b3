#!/bin/bash
# Round 2, lease 6: final-tree validation + ncu evidence (launch list of one step at 1 layer, --set full of the two main kernels
# and of the HBM-bound LoRA kernels).  Every ncu run is preceded by the same command without ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease6; mkdir -p $O
S() { echo "$@" | tee -a $O/summary.txt; }
timeout 500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; S "pytest rc=$? $(tail -1 $O/pytest.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; S "smoke rc=$? $(tail -1 $O/smoke.log)"
timeout 200 python tools/dx_check.py 3 > $O/dx.log 2>&1; S "dx_check rc=$? $(tail -1 $O/dx.log)"
for i in 1 2; do
timeout 240 python bench.py --steps 20 --warmup 5 > $O/bench_$i.out 2> $O/bench_$i.err; S "bench $i rc=$? $(grep -o '"value": [0-9.]*' $O/bench_$i.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench_$i.out | head -1) $(grep -o '"achieved": [0-9.]*' $O/bench_$i.out | head -1)"
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.out 2> $O/bench_ref.err; S "reference arm rc=$? $(grep -o '"value": [0-9.]*' $O/bench_ref.out | head -1)"
timeout 300 python tests/gpu_step_breakdown.py > $O/breakdown.txt 2>&1; cat $O/breakdown.txt | tee -a $O/summary.txt
B="python bench.py --layers 1 --steps 1 --warmup 3 --no-e2e --no-cpu --no-opt"
timeout 200 $B > $O/b_l1.json 2> $O/b_l1.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_step.csv $B > $O/ncu_l1.log 2>&1; S "ncu list rc=$?"
for which in fwd dx; do
  K="python tests/gpu_one_kernel.py $which 5 16384 4096 4096 lora"
  timeout 200 $K > $O/one_$which.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlora_gemm --launch-skip 3 --launch-count 1 -f -o $O/full_$which $K > $O/ncu_full_$which.log 2>&1; S "ncu full $which rc=$?"
done
K="python tests/gpu_skinny_once.py"
timeout 200 $K > $O/skinny.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:qlora_gemm --launch-skip 12 --launch-count 6 -f -o $O/full_skinny $K > $O/ncu_full_skinny.log 2>&1; S "ncu full skinny rc=$?"
