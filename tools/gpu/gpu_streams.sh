#!/bin/bash
# resident GD steps issued on 1 / 2 / 3 caller streams (lane rotation of the _dev entry points)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/streams.txt; : > $O
for s in 1 2 3; do
echo "== --streams $s" >> $O
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --streams $s 2>>gpurun_out/streams.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('gd value %.4e e2e %.4e ms/step %.3f kernel ms %.4f frac %.3f share %s valid %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['roofline']['stage_share_of_step'],d['path_stats']['valid_last_step']))" >> $O
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q 2>&1 | tail -3 >> $O
cat $O; tail -5 gpurun_out/streams.err
