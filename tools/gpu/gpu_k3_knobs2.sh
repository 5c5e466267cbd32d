#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/k3_knobs2.txt; : > $O
for DG in 0 2 3 5; do for DT in 24 36; do
  echo "== CKC_DEEP_GAP=$DG CKC_DEFER_TRI=$DT" >> $O
  CKC_DEEP_GAP=$DG CKC_DEFER_TRI=$DT timeout 900 python tools/k3_probe.py CKC_DEFER_ROUNDS=64,96,128 2>&1 | grep -v create >> $O
done; done
cat $O
