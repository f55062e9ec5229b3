#!/bin/bash
# Development iteration on one B200: GPU tests, sustained per-kernel timings for the given variants, one bench line.
set -u
VARS=${1:-3,4,5}
EXTRA=${2:-}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
rm -f gpurun_out/sustained_bench.jsonl
timeout 900 python tests/gpu_sustained_bench.py --variants $VARS $EXTRA > gpurun_out/sb.log 2>&1; echo "sb rc=$?"
timeout 900 python bench.py --no-cpu > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"
