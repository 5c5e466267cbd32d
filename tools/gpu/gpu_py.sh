#!/bin/bash
# run one python tool on the GPU box: gpu_py.sh tools/x.py [args]; output -> gpurun_out/py.txt
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python "$@" > gpurun_out/py.txt 2>&1; echo "rc=$?" >> gpurun_out/py.txt; tail -40 gpurun_out/py.txt
