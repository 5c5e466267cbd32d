#!/bin/bash
# final single-GPU record of the round: tests, smoke, the three workloads with both arms
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_final_gd.json 2> gpurun_out/bench_gd.err; echo "gd rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_final_gd_ref.json 2> gpurun_out/bench_gd_ref.err; echo "gd ref rc=$?"
python bench.py --workload pcs --steps 10 --warmup 3 --no-extra > gpurun_out/r1_bench_final_pcs.json 2> gpurun_out/bench_pcs.err; echo "pcs rc=$?"
python bench.py --workload pcs --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_final_pcs_ref.json 2> gpurun_out/bench_pcs_ref.err; echo "pcs ref rc=$?"
python bench.py --workload rbs --steps 5 --warmup 3 --no-extra > gpurun_out/r1_bench_final_rbs.json 2> gpurun_out/bench_rbs.err; echo "rbs rc=$?"
python bench.py --workload rbs --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_final_rbs_ref.json 2> gpurun_out/bench_rbs_ref.err; echo "rbs ref rc=$?"
python - <<'PY'
import json
for w in ("gd","pcs","rbs"):
    for arm in ("","_ref"):
        try:
            d=json.loads(open('gpurun_out/r1_bench_final_%s%s.json'%(w,arm)).read().strip().split('\n')[-1])
            print(w,arm,'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'], 'ms/step %.3f'%d['ms_per_step'], 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'))
        except Exception as e: print(w,arm,'ERR',e)
PY
tail -2 gpurun_out/bench_gd.err
