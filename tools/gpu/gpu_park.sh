#!/bin/bash
# K5 straggler parking: exactness test, then K5 time / value / e2e for a few thresholds
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/park.txt; : > $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "gd" 2>&1 | tail -5 >> $O
for p in 0 12 14 16 20; do
echo "== CKC_GD_PARK=$p" >> $O
CKC_GD_PARK=$p timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra 2>>gpurun_out/park.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('gd value %.4e e2e %.4e ms/step %.3f kernel ms %.4f frac %.3f share %s launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['roofline']['stage_share_of_step'],d['gpu_launches']))" >> $O
done
CKC_TIMELINE=1 CKC_GD_PARK=16 python tools/e2e_probe.py gd >> $O 2>&1
CKC_TIMELINE=1 CKC_GD_PARK=16 CKC_PIPE_TAPER=2 CKC_HOST_LANES=8 python tools/e2e_probe.py gd >> $O 2>&1
CKC_GD_PARK=12 python tools/e2e_probe.py gd >> $O 2>&1
cat $O
