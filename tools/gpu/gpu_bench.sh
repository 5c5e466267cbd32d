#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_gd.json 2> gpurun_out/bench_gd.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench_gd.json; tail -5 gpurun_out/bench_gd.err
python bench.py --workload pcs --steps 10 --warmup 3 --no-extra > gpurun_out/bench_pcs.json 2> gpurun_out/bench_pcs.err; echo "rc=$?"; tail -c 2500 gpurun_out/bench_pcs.json; tail -5 gpurun_out/bench_pcs.err
