#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/r1_bench_gd_n$N.json 2> gpurun_out/bench_gd_n$N.err; echo "gd N=$N rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1_bench_gd_n8.json').read().strip().split('\n')[-1])
print('gd_n8 value %.3e e2e %.3e ms/step %.3f n_gpus %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['n_gpus']))
PY
