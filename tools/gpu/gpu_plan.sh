#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export CKC_MESH_PATH=abb_dualarm_planning_b200/data/meshes.ckcmesh
for m in 200 1000 4000; do
  timeout 300 abb_dualarm_planning_b200/host/prm_pcs_batch 1 $m 11 gpurun_out/prm_$m.txt | tail -1
done
timeout 600 python -m pytest tests/test_gpu_planner.py -x -q 2>&1 | tail -8
