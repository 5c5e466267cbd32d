#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-extra --no-cpu"
$CMD > gpurun_out/plain_gd2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gd_project -s 1 -c 1 -o gpurun_out/r1_prof_gd_project_v3 $CMD > gpurun_out/ncu_gd2.log 2>&1
echo "gd full rc=$?"
