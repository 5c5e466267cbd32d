#!/bin/bash
# round 2, experiment 1: fused K3 (frames inside k_collide) parity + timing; GD kernel variants (rolled / unrolled / merged joints)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/exp1.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> gpurun_out/exp1.txt
echo "== K3 probe (fused frames)" >> gpurun_out/exp1.txt
timeout 600 python tools/k3_probe.py CKC_WIDE_CNT 8 2>&1 | tail -4 >> gpurun_out/exp1.txt
for lib in "" abb_dualarm_planning_b200/csrc/variants/libckc_u5.so; do
for v in 0 1 2; do
echo "== GD variant $v lib=${lib:-default(u4)}" >> gpurun_out/exp1.txt
CKC_LIB=${lib:+$GRAFT_REPO_ROOT/$lib} CKC_GD_VARIANT=$v timeout 600 python bench.py --workload gd --steps 20 --warmup 3 --no-cpu --no-extra 2>>gpurun_out/exp1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('value %.4e e2e %.4e ms/step %.3f kernel ms %s frac %.3f valid_last %s iters %.4f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline'].get('kernel_ms_per_launch'),d['roofline']['frac'],d['path_stats']['valid_last_step'],d['path_stats']['newton_iters_mean']))" >> gpurun_out/exp1.txt
done
done
cat gpurun_out/exp1.txt
