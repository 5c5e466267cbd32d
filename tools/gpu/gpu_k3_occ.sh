#!/bin/bash
# K3 against the number of resident warps per SM (same kernel, fewer resident blocks): how much do co-resident warps cost each other?
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/k3_occ.txt; : > $O
echo "== 8-warp blocks: 1, 2 blocks per SM (8, 16 warps)" >> $O
timeout 600 python tools/k3_probe.py CKC_K3_BLOCKS_PER_SM=1,2 2>&1 | grep -v create >> $O
echo "== 4-warp blocks: 1, 2, 3, 4 blocks per SM (4, 8, 12, 16 warps)" >> $O
CKC_LIB=$GRAFT_REPO_ROOT/abb_dualarm_planning_b200/csrc/variants/libckc_w4.so timeout 600 python tools/k3_probe.py CKC_K3_BLOCKS_PER_SM=1,2,3,4 2>&1 | grep -v create >> $O
cat $O
