#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export CKC_MESH_PATH=abb_dualarm_planning_b200/data/meshes.ckcmesh
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for w in 1 4; do
  timeout 600 abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 500 1.0 7 gpurun_out/plan_500_w$w.txt $w | tail -4
done
python bench.py --workload rbs --steps 5 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('RBS value %.4e e2e %.4e'%(d['value'],d['e2e']['value']))"
