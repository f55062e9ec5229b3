#!/bin/bash
# Round 2, lease 20: ncu evidence of the final tree (launch list of one step at 1 layer; --set full of the two main kernels).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease20; mkdir -p $O
B="python bench.py --layers 1 --steps 1 --warmup 3 --no-e2e --no-cpu --no-opt"
timeout 200 $B > $O/b_l1.json 2> $O/b_l1.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_step.csv $B > $O/ncu_l1.log 2>&1; echo "ncu list rc=$?"
for which in fwd dx; do
  K="python tests/gpu_one_kernel.py $which 5 16384 4096 4096 lora"
  timeout 200 $K > $O/one_$which.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlora_gemm --launch-skip 3 --launch-count 1 -f -o $O/full_$which $K > $O/ncu_full_$which.log 2>&1; echo "ncu full $which rc=$?"
done
