#!/bin/bash
# Round 2, lease 17: DRAM traffic of the forward main kernel against the L2 slab size of its rasterisation (ncu, two metrics),
# and the step time for the same settings.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease17; mkdir -p $O
S() { echo "$@" | tee -a $O/summary.txt; }
for mb in 8 16 32 64; do
  for which in fwd dx; do
    K="python tests/gpu_one_kernel.py $which 5 16384 4096 4096 lora"
    B2Q_SLAB_MB=$mb timeout 120 $K > $O/one_${which}_$mb.log 2>&1 && \
    B2Q_SLAB_MB=$mb timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:qlora_gemm --launch-skip 3 --launch-count 1 --csv --log-file $O/ncu_${which}_$mb.csv $K > $O/ncu_${which}_$mb.log 2>&1
    S "slab ${mb} MB $which: $(grep -E 'dram__bytes_read|dram__bytes_write|gpu__time_duration|hit_rate' $O/ncu_${which}_$mb.csv | awk -F'","' '{gsub(/"/,"",$NF); print $(NF-2), $(NF-1), $NF}' | tr '\n' ';')"
  done
done
for rep in 1 2; do for mb in 16 32 64; do
  B2Q_SLAB_MB=$mb timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-opt --no-e2e > $O/ab_slab${mb}_$rep.out 2> $O/ab_slab${mb}_$rep.err
  S "A/B slab=${mb}MB $rep rc=$? $(grep -o '"value": [0-9.]*' $O/ab_slab${mb}_$rep.out | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/ab_slab${mb}_$rep.out | head -1)"
done; done
