#!/bin/bash
# the GPU test suite (+ optional pytest arguments), output -> gpurun_out/tests.txt
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x "$@" 2>&1 | tail -40 > gpurun_out/tests.txt; cat gpurun_out/tests.txt
