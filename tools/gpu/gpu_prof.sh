#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-extra --no-cpu"
$CMD > gpurun_out/plain_gd.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_gd.csv $CMD > gpurun_out/ncu_gd.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_gd2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gd_project -s 1 -c 1 -o gpurun_out/r1_prof_gd_project_final $CMD > gpurun_out/ncu_gd2.log 2>&1
echo "gd full rc=$?"
$CMD > gpurun_out/plain_gd3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_collide -s 2 -c 1 -o gpurun_out/r1_prof_collide_final $CMD > gpurun_out/ncu_gd3.log 2>&1
echo "collide full rc=$?"
CMD2="python bench.py --workload pcs --steps 2 --warmup 1 --no-extra --no-cpu"
$CMD2 > gpurun_out/plain_pcs2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_pcs.csv $CMD2 > gpurun_out/ncu_pcs2.log 2>&1
echo "pcs launch list rc=$?"
$CMD2 > gpurun_out/plain_pcs3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pcs_project -s 1 -c 1 -o gpurun_out/r1_prof_pcs_project $CMD2 > gpurun_out/ncu_pcs3.log 2>&1
echo "pcs project full rc=$?"
