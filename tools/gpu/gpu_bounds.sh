#!/bin/bash
# the -DCKC_BOUNDS build (device-side checks of every stack / queue write and node / triangle index; a violation traps the kernel) over
# the probe sets and the parity / property suites -- the pool has no compute-sanitizer
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/bounds.txt; : > $O
export CKC_LIB=$GRAFT_REPO_ROOT/abb_dualarm_planning_b200/csrc/variants/libckc_bounds.so
timeout 600 python tools/k3_probe.py CKC_WIDE_CNT 8 2>&1 | tail -3 >> $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q -k "not library_is_native and not straggler" 2>&1 | tail -3 >> $O
cat $O
