cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_lease21; mkdir -p $O
for shape in "16384 14336 4096" "16384 4096 14336"; do
  tag=$(echo $shape | tr ' ' 'x')
  for mb in 8 16 32; do
    for which in fwd dx; do
      K="python tests/gpu_one_kernel.py $which 5 $shape lora"
      B2Q_SLAB_MB=$mb timeout 120 $K > $O/one_${which}_${tag}_$mb.log 2>&1 && \
      B2Q_SLAB_MB=$mb timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:qlora_gemm --launch-skip 3 --launch-count 1 --csv --log-file $O/ncu_${which}_${tag}_$mb.csv $K > $O/ncu_${which}_${tag}_$mb.log 2>&1
      echo "M,N,K=$shape slab ${mb} MB $which: $(grep -E 'dram__bytes_read|dram__bytes_write|gpu__time_duration|hit_rate' $O/ncu_${which}_${tag}_$mb.csv | awk -F'","' '{gsub(/"/,"",$NF); print $(NF-2), $NF}' | tr '\n' ';')" | tee -a $O/summary.txt
    done
  done
done
