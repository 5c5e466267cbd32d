#!/bin/bash
# candidate K3 thresholds on the bench workloads (same box): GD step (collide share x ms/step), PCS, RBS
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/k3_knobs3.txt; : > $O
for cfg in "CKC_DEFER_ROUNDS=64 CKC_DEEP_GAP=0" "CKC_DEFER_ROUNDS=128 CKC_DEEP_GAP=0" "CKC_DEFER_ROUNDS=128 CKC_DEEP_GAP=3"; do
echo "== $cfg" >> $O
for wl in gd pcs rbs; do
env $cfg timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-extra 2>>gpurun_out/k3_knobs3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); r=d['roofline']
print('$wl value %.4e ms/step %.3f collide ms/launch %.4f share %s'%(d['value'],d['ms_per_step'],r['kernel_ms_per_launch'] if '$wl'!='gd' else r['stage_share_of_step']['collide']*float(r['timing_note'].split('stream, ')[1].split(' ms')[0]),r['stage_share_of_step']))" >> $O
done; done
cat $O
