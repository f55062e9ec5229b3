#!/bin/bash
# Round 2, lease 14: L2 persistence for dx between the decode GEMM and the masked add (B2Q_DX_L2_PERSIST=<MB>), same-box A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease14; mkdir -p $O
S() { echo "$@" | tee -a $O/summary.txt; }
python - <<'PY' | tee -a $O/summary.txt
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size >> 20, "MB; persisting max", getattr(p, "persisting_l2_cache_max_size", -1) >> 20, "MB; window max", getattr(p, "access_policy_max_window_size", -1) >> 20, "MB")
PY
B2Q_DX_L2_PERSIST=64 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dropout or full_size or c1 or repeated" > $O/pytest_persist.log 2>&1; S "pytest persist=64 rc=$? $(tail -1 $O/pytest_persist.log)"
for rep in 1 2; do for v in 0 32 64 96; do
  B2Q_DX_L2_PERSIST=$v timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_p${v}_$rep.out 2> $O/ab_p${v}_$rep.err
  S "A/B persist=${v}MB $rep rc=$? $(grep -o '"value": [0-9.]*' $O/ab_p${v}_$rep.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/ab_p${v}_$rep.out | head -1)"
done; done
for v in 0 64; do B2Q_DX_L2_PERSIST=$v timeout 300 python tests/gpu_step_breakdown.py > $O/breakdown_p$v.txt 2>&1; S "persist=$v $(grep -E 'step|qlora_bwd_dx|qlora_fwd' $O/breakdown_p$v.txt | tr '\n' ' ')"; done
