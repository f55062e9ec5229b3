#!/bin/bash
# Round 2, lease 18: step time against the L2 slab size (8 / 12 / 16 / 24 MB), same box.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease18; mkdir -p $O
S() { echo "$@" | tee -a $O/summary.txt; }
for rep in 1 2; do for mb in 8 12 16 24; do
  B2Q_SLAB_MB=$mb timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_slab${mb}_$rep.out 2> $O/ab_slab${mb}_$rep.err
  S "A/B slab=${mb}MB $rep rc=$? $(grep -o '"value": [0-9.]*' $O/ab_slab${mb}_$rep.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/ab_slab${mb}_$rep.out | head -1)"
done; done
for mb in 8 16; do B2Q_SLAB_MB=$mb timeout 300 python tests/gpu_step_breakdown.py > $O/breakdown_$mb.txt 2>&1; S "slab=$mb $(grep -E 'step|qlora_fwd|qlora_bwd_dx' $O/breakdown_$mb.txt | tr -s ' ' | tr '\n' ' ')"; done
