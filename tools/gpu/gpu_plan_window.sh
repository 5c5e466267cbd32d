#!/bin/bash
# batching inside one query: planner tests, then single-query latency / 500-query throughput for the window sizes; then the K3 captures
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/plan_window.txt; : > $O
timeout 900 python -m pytest tests/test_gpu_planner.py -x -q 2>&1 | tail -5 >> $O
export CKC_MESH_PATH=abb_dualarm_planning_b200/data/meshes.ckcmesh
for w in 1 4 16 64; do
echo "== window $w: 40 queries one after another" >> $O
CKC_PLAN_WINDOW=$w abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 40 1.0 7 /tmp/p1_$w.txt 0 pcs | tail -2 >> $O
tail -1 /tmp/p1_$w.txt | tr ' ' '\n' | paste - - | grep -E "latency|gpu_calls|mean_rounds" | tr '\n' ' ' >> $O; echo >> $O
echo "== window $w: 500 queries, 4 workers" >> $O
CKC_PLAN_WINDOW=$w abb_dualarm_planning_b200/host/cbirrt_pcs_batch 1 500 1.0 7 /tmp/p500_$w.txt 4 pcs | tail -1 >> $O
done
cmp <(grep -v summary /tmp/p1_1.txt) <(grep -v summary /tmp/p1_16.txt) && echo "paths window 1 == window 16: identical" >> $O
cmp <(grep -v summary /tmp/p500_1.txt) <(grep -v summary /tmp/p500_64.txt) && echo "500 paths window 1 == window 64: identical" >> $O
cat $O
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:"^k_collide$" -s 4 -c 1 -o gpurun_out/r2_prof_collide $CMD > gpurun_out/ncu_gd3.log 2>&1; echo "collide full rc=$?"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"^k_collide$" -s 4 -c 1 -o gpurun_out/r2_prof_collide_warm $CMD > gpurun_out/ncu_gd5.log 2>&1; echo "collide warm rc=$?"
CMD2="python bench.py --workload pcs --steps 2 --warmup 3 --no-extra --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:"^k_collide$" -s 6 -c 1 -o gpurun_out/r2_prof_collide_pcs $CMD2 > gpurun_out/ncu_pcs2.log 2>&1; echo "pcs collide full rc=$?"
ls -la gpurun_out/*.ncu-rep
