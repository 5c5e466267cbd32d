#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/quick.txt
for wl in gd pcs; do
python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('$wl value %.4e e2e %.4e ms/step %.3f kernel ms %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline'].get('kernel_ms_per_launch')))" >> gpurun_out/quick.txt
done
cat gpurun_out/quick.txt
