#!/bin/bash
# K3 change check: probe (times + mask hashes), the collision / mask parity tests, RBS bench
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
tools/gpu/gpu_k3.sh "$@"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python bench.py --workload rbs --steps 5 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('RBS value %.4e'%d['value'], d['roofline']['stage_share_of_step'], d['path_stats'])"
