#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/rbs_tail.txt; : > $O
for t in 0 2048 4096 8192 16384 32768; do
echo "== CKC_RBS_TAIL=$t" >> $O
CKC_RBS_TAIL=$t timeout 600 python bench.py --workload rbs --steps 5 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('rbs value %.4e ms/step %.3f launches %d share %s %s'%(d['value'],d['ms_per_step'],d['gpu_launches'],d['roofline']['stage_share_of_step'],d['path_stats']))" >> $O
done
cat $O
