#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export CKC_MESH_PATH=abb_dualarm_planning_b200/data/meshes.ckcmesh
for i in 1 2 3; do
timeout 600 abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 500 2.6 7 gpurun_out/plan_gd_500.txt 4 gd | tail -1
done
for i in 1 2; do
timeout 600 abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 500 2.6 7 gpurun_out/plan_gd_500.txt 1 gd | tail -1
done
