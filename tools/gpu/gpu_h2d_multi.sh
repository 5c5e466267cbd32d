#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; : > gpurun_out/h2d_multi.txt
for N in 1 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N tools/h2d_multi_probe.py 2>/dev/null | grep H2D >> gpurun_out/h2d_multi.txt
done
nproc >> gpurun_out/h2d_multi.txt; cat gpurun_out/h2d_multi.txt
