#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python bench.py --steps 10 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('GD value %.3e e2e %.3e ms/step %.3f kernel ms %.4f share %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],d['roofline']['stage_share_of_step']))"
