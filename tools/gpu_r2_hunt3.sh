#!/bin/bash
# Round 2, third lease: silent corruption in the decode GEMMs -- which build variants show it, and in which k-block.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2_hunt3; mkdir -p $O
P=$PWD/causal-unified-language-vision_b200
for v in r1 cta_noguard ctascope cta_dep cta_syncw cta_oldroles default; do
  if [ $v = default ]; then unset B2Q_LIB_PATH; else export B2Q_LIB_PATH=$P/libb2q_$v.so; fi
  timeout 200 python tools/kb_probe.py 6 > $O/kb_$v.log 2>&1; echo "kb_probe $v rc=$? $(tail -1 $O/kb_$v.log)" | tee -a $O/summary.txt
  timeout 200 python tools/dx_check.py 2 > $O/dx_$v.log 2>&1; echo "dx_check $v rc=$? $(tail -1 $O/dx_$v.log)" | tee -a $O/summary.txt
done
cat $O/summary.txt
