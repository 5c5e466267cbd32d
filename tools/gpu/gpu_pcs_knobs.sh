#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; : > gpurun_out/pcs_knobs.txt
run() { echo "== $*" >> gpurun_out/pcs_knobs.txt; env "$@" timeout 600 python bench.py --workload pcs --steps 20 --warmup 5 --no-cpu --no-extra 2>>gpurun_out/pcs_knobs.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('pcs value %.4e e2e %.4e ms/step %.3f kernel ms %.4f share %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],d['roofline']['stage_share_of_step']))" >> gpurun_out/pcs_knobs.txt; }
run CKC_OVERLAP=1
run CKC_OVERLAP=0
run CKC_OVERLAP=1 CKC_DEFER_ROUNDS=48 CKC_DEFER_TRI=16
run CKC_OVERLAP=1 CKC_DEFER_ROUNDS=32 CKC_DEFER_TRI=12
cat gpurun_out/pcs_knobs.txt
