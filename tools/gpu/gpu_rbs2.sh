#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; : > gpurun_out/rbs2.txt
for tail in 4096 0 4096 0; do
echo "== CKC_RBS_TAIL=$tail" >> gpurun_out/rbs2.txt
CKC_RBS_TAIL=$tail timeout 600 python bench.py --workload rbs --steps 5 --warmup 2 --no-cpu --no-extra 2>>gpurun_out/rbs2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('rbs value %.4e ms/step %.3f R %d checks/edge %.2f launches %d share %s levels %.1f'%(d['value'],d['ms_per_step'],d['config']['inner_repeats'],d['path_stats']['checks_per_edge'],d['gpu_launches'],d['roofline']['stage_share_of_step'],d['path_stats']['levels_per_call']))" >> gpurun_out/rbs2.txt
done
cat gpurun_out/rbs2.txt; tail -3 gpurun_out/rbs2.err
