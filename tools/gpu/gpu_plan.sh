#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export CKC_MESH_PATH=abb_dualarm_planning_b200/data/meshes.ckcmesh
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 500 1.0 7 gpurun_out/plan_500.txt 4 pcs | tail -2
timeout 600 abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 500 2.6 7 gpurun_out/plan_gd_500.txt 4 gd | tail -2
python bench.py --workload rbs --steps 5 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('RBS value %.4e'%d['value'], d['roofline']['stage_share_of_step'])"
