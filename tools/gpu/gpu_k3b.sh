#!/bin/bash
# K3 with the heavy routine fused into the end of the launch: masks, timing, threshold sweep
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; : > gpurun_out/k3b.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q 2>&1 | tail -4 >> gpurun_out/k3b.txt
for th in "64 24" "32 8" "24 6" "16 4" "48 16" "1000000 1000000"; do set -- $th
echo "== CKC_DEFER_ROUNDS=$1 CKC_DEFER_TRI=$2" >> gpurun_out/k3b.txt
CKC_DEFER_ROUNDS=$1 CKC_DEFER_TRI=$2 timeout 300 python tools/k3_probe.py CKC_WIDE_CNT 8 2>&1 | tail -3 | cut -c17-150 >> gpurun_out/k3b.txt
done
cat gpurun_out/k3b.txt
