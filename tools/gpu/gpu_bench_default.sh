#!/bin/bash
# the driver's command: default bench (GD headline + PCS / RBS sub-records + planning extras), then the reference arm
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"
tail -5 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().split('\n')[-1])
def show(name, r):
    print('%-4s value %.4e e2e %.4e ms/step %.3f region %.2fs R %s | roofline %s frac %.3f kernel ms %.4f | cpu %s' % (name, r['value'], r['e2e']['value'], r['ms_per_step'],
        r.get('timed_region_s', r.get('config',{}).get('timed_region_s', 0)) or 0, r.get('inner_repeats', r.get('config',{}).get('inner_repeats')), r['roofline']['bound'], r['roofline']['frac'], r['roofline']['kernel_ms_per_launch'],
        (r.get('cpu_baseline') or {}).get('value')))
show('gd', d); show('pcs', d['extra']['pcs']); show('rbs', d['extra']['rbs'])
print(json.dumps(d['extra'].get('planning'), indent=1)[:1500])
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
print(d['roofline']['stage_share_of_step'], d['extra']['pcs']['roofline']['stage_share_of_step'], d['extra']['rbs']['roofline']['stage_share_of_step'])
print(d['extra']['rbs']['roofline']['timing_note'])
PY
