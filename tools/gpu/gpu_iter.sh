#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_gd.json 2> gpurun_out/bench_gd.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_gd.json').read().strip().split('\n')[-1])
print('GD value %.3e e2e %.3e ms/step %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step']))
print(d['roofline']['stage_share_of_step'], 'kernel ms', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'])
print(d.get('extra'))
PY
tail -3 gpurun_out/bench_gd.err
