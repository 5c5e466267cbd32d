#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for f in 0 1; do
CKC_GD_FAN=$f python bench.py --steps 10 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('fan=$f GD value %.3e e2e %.3e ms/step %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step']))"
done
