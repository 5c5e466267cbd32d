#!/bin/bash
# round 2 profiling evidence: launch lists (GD default workload, PCS, RBS) and `ncu --set full` captures of the dominant kernels.
# Every ncu command runs only after the identical plain command exited 0.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; rm -f gpurun_out/r2_prof_*.ncu-rep
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu"
$CMD > gpurun_out/plain_gd.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_gd.csv $CMD > gpurun_out/ncu_gd.log 2>&1
echo "gd launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_gd_project -s 3 -c 1 -o gpurun_out/r2_prof_gd_project $CMD > gpurun_out/ncu_gd2.log 2>&1
echo "gd full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"^k_collide$" -s 4 -c 1 -o gpurun_out/r2_prof_collide $CMD > gpurun_out/ncu_gd3.log 2>&1
echo "collide full rc=$?"
CMD2="python bench.py --workload pcs --steps 2 --warmup 3 --no-extra --no-cpu"
$CMD2 > gpurun_out/plain_pcs.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_pcs.csv $CMD2 > gpurun_out/ncu_pcs.log 2>&1
echo "pcs launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"^k_collide$" -s 6 -c 1 -o gpurun_out/r2_prof_collide_pcs $CMD2 > gpurun_out/ncu_pcs2.log 2>&1
echo "pcs collide full rc=$?"
CMD3="python bench.py --workload rbs --steps 1 --warmup 1 --samples 20000 --no-extra --no-cpu"
$CMD3 > gpurun_out/plain_rbs.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_rbs.csv $CMD3 > gpurun_out/ncu_rbs.log 2>&1
echo "rbs launch list rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_*.csv
