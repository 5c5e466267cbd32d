#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu"
$CMD > gpurun_out/plain_gd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_collide -s 4 -c 1 -o gpurun_out/r2_prof_collide $CMD > gpurun_out/ncu_gd3.log 2>&1
echo "collide full rc=$?"; tail -3 gpurun_out/ncu_gd3.log
