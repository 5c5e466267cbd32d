#!/bin/bash
# resident PCS steps: fan-out over the lanes inside a call (CKC_OVERLAP=1) vs whole-batch launches (0), on 1 / 2 caller streams
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/pcs_streams.txt; : > $O
for ov in 1 0; do for s in 1 2; do
echo "== CKC_OVERLAP=$ov --streams $s" >> $O
CKC_OVERLAP=$ov timeout 600 python bench.py --workload pcs --steps 20 --warmup 5 --no-cpu --no-extra --streams $s 2>>gpurun_out/pcs_streams.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('pcs value %.4e e2e %.4e ms/step %.3f kernel ms %.4f share %s valid %s one-stream %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],d['roofline']['stage_share_of_step'],d['path_stats']['valid_last_step'] if d.get('path_stats') else None, d.get('one_caller_stream')))" >> $O
done; done
cat $O; tail -3 gpurun_out/pcs_streams.err
