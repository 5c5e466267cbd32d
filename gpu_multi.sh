#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_gd_n2.json 2> gpurun_out/bench_gd_n2.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_gd_n2.json; tail -3 gpurun_out/bench_gd_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference 2>&1 | tail -2 | cut -c1-600
