#!/bin/bash
# K3 probe + the GPU tests that pin the collision masks
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python tools/k3_probe.py CKC_WIDE_CNT 8 2>&1 | tail -3 > gpurun_out/k3.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "collision or mask or nan or env2 or isvalid or is_valid or rbs" 2>&1 | tail -4 >> gpurun_out/k3.txt
cat gpurun_out/k3.txt
