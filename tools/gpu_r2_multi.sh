#!/bin/bash
# Multi-GPU lease: data-parallel parity (both gradient-sync backends) and the bench exactly as the driver launches it.
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r2_multi.sh N REPS [dp]'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}; REPS=${2:-1}; DP=${3:-}
O=gpurun_out/r2_multi_n$N; mkdir -p $O
PORT=29600
S() { echo "$@" | tee -a $O/summary.txt; }
nvidia-smi topo -m > $O/topo.txt 2>&1
if [ -n "$DP" ]; then
  for be in torch b2q; do
    B2Q_COMM_BACKEND=$be timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port $PORT tests/gpu_dp_check.py > $O/dp_$be.log 2>&1
    S "dp_check backend=$be rc=$? $(tail -1 $O/dp_$be.log | cut -c1-120)"; PORT=$((PORT + 1))
  done
fi
for i in $(seq 1 $REPS); do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_$i.out 2> $O/bench_$i.err
  S "bench N=$N run $i rc=$? $(grep -o '"value": [0-9.]*' $O/bench_$i.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/bench_$i.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench_$i.out | head -1)"
  PORT=$((PORT + 1))
done
grep -h "stall guard" -A 6 $O/*.err | head -60 > $O/stalls.txt
