#!/bin/bash
# Round evidence run on one B200: GPU tests, bench line, ncu launch list of one step, ncu --set full of the two main kernels.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
B="python bench.py --layers 1 --steps 1 --warmup 3 --no-e2e --no-cpu --no-opt"
$B > gpurun_out/b_l1.json 2> gpurun_out/b_l1.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_step.csv $B > gpurun_out/ncu_l1.log 2>&1; echo "ncu list rc=$?"
for which in fwd dx; do
  K="python tests/gpu_one_kernel.py $which 5 16384 4096 4096 lora"
  $K > gpurun_out/one_$which.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:qlora_gemm --launch-skip 3 --launch-count 1 -f -o gpurun_out/full_$which $K > gpurun_out/ncu_full_$which.log 2>&1; echo "ncu full $which rc=$?"
done
