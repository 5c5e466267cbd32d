#!/bin/bash
# what the driver runs at round end, in its order: GPU test suite, smoke(), default bench, reference arm
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
tools/gpu/gpu_tests.sh
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
tools/gpu/gpu_bench_default.sh
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/bench_default_ref.json | cut -c1-300
