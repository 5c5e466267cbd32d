#!/bin/bash
# K3: sweep of the deferral thresholds (box rounds / triangle steps after which a configuration goes to k_collide_heavy)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; : > gpurun_out/k3sweep.txt
for dr in 16 24 32; do for dt in 2 4 8; do
echo "== CKC_DEFER_ROUNDS=$dr CKC_DEFER_TRI=$dt" >> gpurun_out/k3sweep.txt
CKC_DEFER_ROUNDS=$dr CKC_DEFER_TRI=$dt timeout 300 python tools/k3_probe.py CKC_WIDE_CNT 8 2>&1 | tail -3 | cut -c17-150 >> gpurun_out/k3sweep.txt
done; done
cat gpurun_out/k3sweep.txt
