#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 800 python tools/split_probe.py > gpurun_out/split.txt 2>&1; cat gpurun_out/split.txt | tail -20
