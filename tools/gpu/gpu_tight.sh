#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/tight.txt; : > $O
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4 >> $O
timeout 900 python tools/gd_agreement_probe.py >> $O 2>&1
cat $O
