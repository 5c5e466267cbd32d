#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export CKC_MESH_PATH=abb_dualarm_planning_b200/data/meshes.ckcmesh
timeout 900 python -m pytest tests/test_gpu_planner.py -x -q 2>&1 | tail -12
for st in 1.0 2.6; do
  timeout 600 abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 500 $st 7 gpurun_out/plan_gd_500_$st.txt 4 gd | tail -4
done
