#!/bin/bash
# round 2, experiment 2: K3 shared-memory footprint (L1 carve-out) A/B, ncu captures of the fused K3 and the unrolled GD kernel
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
echo "== K3 probe, stack 512 (46 kB / block)" > gpurun_out/exp2.txt
timeout 600 python tools/k3_probe.py CKC_WIDE_CNT 8 2>&1 | tail -3 >> gpurun_out/exp2.txt
echo "== K3 probe, stack 640 (50 kB / block)" >> gpurun_out/exp2.txt
CKC_LIB=$GRAFT_REPO_ROOT/abb_dualarm_planning_b200/csrc/variants/libckc_s640.so timeout 600 python tools/k3_probe.py CKC_WIDE_CNT 8 2>&1 | tail -3 >> gpurun_out/exp2.txt
CMD="python bench.py --steps 2 --warmup 1 --no-extra --no-cpu"
$CMD > gpurun_out/plain_gd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gd_project -s 1 -c 1 -o gpurun_out/r2_prof_gd_project_u $CMD > gpurun_out/ncu_gd2.log 2>&1
echo "gd full rc=$?" >> gpurun_out/exp2.txt
ncu --set full --clock-control none --import-source on -k regex:k_collide -s 2 -c 1 -o gpurun_out/r2_prof_collide_fused $CMD > gpurun_out/ncu_gd3.log 2>&1
echo "collide full rc=$?" >> gpurun_out/exp2.txt
cat gpurun_out/exp2.txt
