#!/bin/bash
# Multi-GPU soak: the bench with 100 timed steps per phase (100 C-ABI steps + 100 module-surface steps per rank).
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r2_soak_multi.sh N'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-4}
O=gpurun_out/r2_soak_n$N; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 \
    bench.py --gpus $N --steps ${2:-100} --warmup 5 --e2e-timeout 300 --global-timeout 560 > $O/bench.out 2> $O/bench.err
echo "soak N=$N rc=$? $(grep -o '"value": [0-9.]*' $O/bench.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/bench.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench.out | head -1)" | tee $O/summary.txt
grep -h "stall guard" -A 6 $O/bench.err | head -40 >> $O/summary.txt
