#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for wl in gd pcs rbs; do
python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('$wl value %.4e e2e %.4e ms/step %.3f kernel ms %.4f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch']))"
done
