#!/bin/bash
# K3 probe only (time + counters + mask hash on the 94 k / 45 k / 31.5 k configuration sets); optional CKC_LIB variants as arguments
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
: > gpurun_out/k3.txt
for lib in "" "$@"; do
echo "== lib=${lib:-default}" >> gpurun_out/k3.txt
CKC_LIB=${lib:+$GRAFT_REPO_ROOT/$lib} timeout 600 python tools/k3_probe.py CKC_WIDE_CNT 8 2>&1 | tail -3 >> gpurun_out/k3.txt
done
cat gpurun_out/k3.txt
