#!/bin/bash
# Round 2, lease 8: masked dX GEMM walking the decode GEMM's L2 slabs last-in-first-out (B2Q_DX_LIFO=1, default) vs in order.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease8; mkdir -p $O
S() { echo "$@" | tee -a $O/summary.txt; }
timeout 500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; S "pytest rc=$? $(tail -1 $O/pytest.log)"
for rep in 1 2 3; do for v in 0 1; do
  B2Q_DX_LIFO=$v timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_lifo${v}_$rep.out 2> $O/ab_lifo${v}_$rep.err
  S "A/B lifo=$v $rep rc=$? $(grep -o '"value": [0-9.]*' $O/ab_lifo${v}_$rep.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/ab_lifo${v}_$rep.out | head -1)"
done; done
for v in 0 1; do B2Q_DX_LIFO=$v timeout 300 python tests/gpu_step_breakdown.py > $O/breakdown_lifo$v.txt 2>&1; S "lifo=$v $(grep qlora_bwd_dx $O/breakdown_lifo$v.txt)"; done
