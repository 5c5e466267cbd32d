#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD2="python bench.py --workload pcs --steps 2 --warmup 1 --no-extra --no-cpu"
$CMD2 > gpurun_out/plain_pcs.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_collide -s 6 -c 1 -o gpurun_out/r1_prof_collide_v2 $CMD2 > gpurun_out/ncu_pcs.log 2>&1
echo "collide full rc=$?"
