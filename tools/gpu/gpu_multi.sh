#!/bin/bash
# one 8-GPU box: the driver's scaling command at N = 8 (default workload with the PCS / RBS sub-records), then N = 8 again with the host
# side spread differently (CKC_HOST_LANES) for the e2e question, and the in-library multi-device context (ckc_create_multi) from ONE process
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_gd_n$N.json 2> gpurun_out/bench_gd_n$N.err; echo "gd N=$N rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_gd_n$N.json').read().strip().split('\n')[-1])
print('gd N=$N value %.3e e2e %.3e ms/step %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d['config']['host_binding_rank0'])
for k in ('pcs','rbs'):
    r=d['extra'][k]; print(k,'value %.3e e2e %.3e ms/step %.3f'%(r['value'],r['e2e']['value'],r['ms_per_step']))
PY
[ "$N" = "8" ] && { python tools/multi_probe.py $N > gpurun_out/multi_probe.txt 2>&1; cat gpurun_out/multi_probe.txt; } || true
