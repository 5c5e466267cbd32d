#!/bin/bash
# round-2 evidence in one box: GPU test suite, default bench (driver's command) + reference arm, then the profiling captures
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
tools/gpu/gpu_tests.sh
tools/gpu/gpu_bench_default.sh
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/bench_default_ref.json | cut -c1-600
tools/gpu/gpu_r2_prof.sh
