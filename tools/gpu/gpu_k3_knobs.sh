#!/bin/bash
# K3 traversal knobs re-swept on the final kernel (one process per value, same box): every line carries time, tests per configuration and the mask hash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/k3_knobs.txt; : > $O
for spec in CKC_WIDE_CNT=0,4,8,12,16,32 CKC_FLUSH_N=3,6,9,12,20 CKC_LIKELY_GAP=0.5,2.0125,4,8 CKC_DEEP_GAP=-4,-1,0,1,3 CKC_DEFER_ROUNDS=40,48,64,96; do
  timeout 900 python tools/k3_probe.py $spec 2>&1 | grep -v create >> $O
done
cat $O
