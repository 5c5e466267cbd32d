#!/bin/bash
# RBS tail kernel: parity tests, the RBS bench line with and without the tail, planner single-query latency
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "rbs or RBS or reconstruct or dropin or members or planner or multi" 2>&1 | tail -15 > gpurun_out/rbs.txt
for tail in 4096; do
echo "== CKC_RBS_TAIL=$tail" >> gpurun_out/rbs.txt
CKC_RBS_TAIL=$tail timeout 600 python bench.py --workload rbs --steps 5 --warmup 2 --no-cpu --no-extra 2>>gpurun_out/rbs.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('rbs value %.4e ms/step %.3f checks/edge %.2f ok %.3f launches %d share %s'%(d['value'],d['ms_per_step'],d['path_stats']['checks_per_edge'],d['path_stats']['edges_ok_fraction'],d['gpu_launches'],d['roofline']['stage_share_of_step']))" >> gpurun_out/rbs.txt
export CKC_MESH_PATH=$GRAFT_REPO_ROOT/abb_dualarm_planning_b200/data/meshes.ckcmesh
CKC_RBS_TAIL=$tail abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 40 1.0 7 gpurun_out/plan_seq_$tail.txt 0 pcs | tail -4 >> gpurun_out/rbs.txt
tail -1 gpurun_out/plan_seq_$tail.txt | cut -c1-400 >> gpurun_out/rbs.txt
done
cat gpurun_out/rbs.txt
