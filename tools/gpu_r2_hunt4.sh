#!/bin/bash
# Round 2, fourth lease: both fixes in (phase aliasing, packed-slot load fence), .cta scope default.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_hunt4; mkdir -p $O
P=$PWD/causal-unified-language-vision_b200
S() { echo "$@" | tee -a $O/summary.txt; }
for i in 1 2; do timeout 500 python -m pytest tests -m gpu -q > $O/pytest_$i.log 2>&1; S "pytest $i rc=$? $(tail -1 $O/pytest_$i.log)"; done
for v in default noguard; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 200 python tools/kb_probe.py 8 > $O/kb_$v.log 2>&1; S "kb_probe $v rc=$? $(tail -1 $O/kb_$v.log)"
  timeout 200 python tools/dx_check.py 3 > $O/dx_$v.log 2>&1; S "dx_check $v rc=$? $(tail -1 $O/dx_$v.log)"
done
unset B2Q_LIB_PATH
for i in 1 2 3; do
  timeout 240 python bench.py --steps 20 --warmup 5 > $O/bench_$i.out 2> $O/bench_$i.err
  S "bench $i rc=$? $(grep -o '"value": [0-9.]*' $O/bench_$i.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench_$i.out | head -1)"
done
for i in 1 2; do timeout 240 python tools/stall_hunt.py --iters 10 > $O/hunt_$i.out 2> $O/hunt_$i.err; S "hunt $i rc=$? $(tail -1 $O/hunt_$i.err | cut -c1-140)"; done
for v in r1 noguard cluster default; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_$v.out 2> $O/ab_$v.err
  S "A/B $v rc=$? $(grep -o '"value": [0-9.]*' $O/ab_$v.out | head -1) $(grep -o '"achieved": [0-9.]*' $O/ab_$v.out | head -1)"
done
unset B2Q_LIB_PATH
B2Q_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_c_host.py -m gpu -x -q > $O/c_host.log 2>&1; S "c host rc=$? $(tail -1 $O/c_host.log)"
B2Q_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variant6 or linear_fwd_bwd or mask_bits" > $O/experimental.log 2>&1; S "experimental parity rc=$? $(tail -1 $O/experimental.log)"
B2Q_FWD_VARIANT=6 B2Q_DX_VARIANT=6 timeout 200 python tools/dx_check.py 3 > $O/dx_v6.log 2>&1; S "dx_check v6 rc=$? $(tail -1 $O/dx_v6.log)"
B2Q_FWD_VARIANT=6 B2Q_DX_VARIANT=6 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_v6.out 2> $O/ab_v6.err; S "A/B v6 rc=$? $(grep -o '"value": [0-9.]*' $O/ab_v6.out | head -1) $(grep -o '"achieved": [0-9.]*' $O/ab_v6.out | head -1)"
B2Q_MASK_BITS=1 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_maskbits.out 2> $O/ab_maskbits.err; S "A/B maskbits rc=$? $(grep -o '"value": [0-9.]*' $O/ab_maskbits.out | head -1)"
B2Q_DX_MASK_FIRST=1 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_maskfirst.out 2> $O/ab_maskfirst.err; S "A/B maskfirst rc=$? $(grep -o '"value": [0-9.]*' $O/ab_maskfirst.out | head -1)"
timeout 120 python tools/stall_selftest.py > $O/selftest.log 2>&1; S "selftest rc=$?"
