#!/bin/bash
# Round 2, lease 10: programmatic dependent launch on/off on the final tree (parity first), same-box A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease10; mkdir -p $O
S() { echo "$@" | tee -a $O/summary.txt; }
B2Q_PDL=1 timeout 500 python -m pytest tests -m gpu -q > $O/pytest_pdl.log 2>&1; S "pytest PDL=1 rc=$? $(tail -1 $O/pytest_pdl.log)"
B2Q_PDL=1 timeout 200 python tools/dx_check.py 2 > $O/dx_pdl.log 2>&1; S "dx_check PDL=1 rc=$? $(tail -1 $O/dx_pdl.log)"
for rep in 1 2 3; do for v in 0 1; do
  B2Q_PDL=$v timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_pdl${v}_$rep.out 2> $O/ab_pdl${v}_$rep.err
  S "A/B PDL=$v $rep rc=$? $(grep -o '"value": [0-9.]*' $O/ab_pdl${v}_$rep.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/ab_pdl${v}_$rep.out | head -1)"
done; done
B2Q_PDL=1 timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_pdl.out 2> $O/bench_pdl.err; S "bench PDL=1 full rc=$? $(grep -o '"value": [0-9.]*' $O/bench_pdl.out | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/bench_pdl.out | head -1)"
B2Q_PDL=1 timeout 300 python tools/stall_hunt.py --iters 20 > $O/hunt_pdl.out 2> $O/hunt_pdl.err; S "hunt PDL=1 rc=$? $(tail -1 $O/hunt_pdl.err | cut -c1-120)"
