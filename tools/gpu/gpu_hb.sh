#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/hb.txt; : > $O
timeout 900 python -m pytest tests/test_gpu_planner.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -5 >> $O
export CKC_MESH_PATH=abb_dualarm_planning_b200/data/meshes.ckcmesh
abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 40 1.0 7 /tmp/hb1.txt 0 hb | tail -4 >> $O
tail -1 /tmp/hb1.txt >> $O
abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 500 1.0 7 /tmp/hb500.txt 4 hb | tail -1 >> $O
cat $O
