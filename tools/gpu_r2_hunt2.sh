#!/bin/bash
# Round 2, second lease: after the phase-aliasing fix -- dX parity repeats (3 builds), tests, bench / hunt loops.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_hunt2; mkdir -p $O
P=$PWD/causal-unified-language-vision_b200
timeout 120 python tools/stall_selftest.py > $O/selftest.log 2>&1; echo "selftest rc=$?" | tee -a $O/summary.txt
for v in default ctascope noguard; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 300 python tools/dx_check.py 4 > $O/dx_$v.log 2>&1; echo "dx_check $v rc=$? $(tail -1 $O/dx_$v.log)" | tee -a $O/summary.txt
done
unset B2Q_LIB_PATH
for i in 1 2 3; do
  timeout 400 python -m pytest tests -m gpu -q > $O/pytest_$i.log 2>&1; echo "pytest $i rc=$? $(tail -1 $O/pytest_$i.log)" | tee -a $O/summary.txt
done
for i in 1 2 3 4 5; do
  timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_$i.out 2> $O/bench_$i.err
  echo "bench $i rc=$? $(grep -o '"value": [0-9.]*' $O/bench_$i.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench_$i.out | head -1)" | tee -a $O/summary.txt
  timeout 240 python tools/stall_hunt.py --iters 10 > $O/hunt_$i.out 2> $O/hunt_$i.err
  echo "hunt $i rc=$? $(tail -1 $O/hunt_$i.err | cut -c1-160)" | tee -a $O/summary.txt
done
export B2Q_LIB_PATH=$P/libb2q_noguard.so
timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_noguard.out 2> $O/bench_noguard.err
echo "bench noguard rc=$? $(grep -o '"value": [0-9.]*' $O/bench_noguard.out | head -1)" | tee -a $O/summary.txt
export B2Q_LIB_PATH=$P/libb2q_ctascope.so
timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_ctascope.out 2> $O/bench_ctascope.err
echo "bench ctascope rc=$? $(grep -o '"value": [0-9.]*' $O/bench_ctascope.out | head -1)" | tee -a $O/summary.txt
cat $O/summary.txt
