#!/bin/bash
# (a) front-tapered chunk schedule of the host pipeline; (b) K3 with child prefetch (variant library)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/exp3.txt; : > $O
for wl in gd pcs; do
CKC_PIPE_TAPER_FRONT=0 python tools/e2e_probe.py $wl >> $O 2>&1
CKC_PIPE_TAPER_FRONT=1 CKC_HOST_LANES=8 python tools/e2e_probe.py $wl >> $O 2>&1
CKC_TIMELINE=1 CKC_PIPE_TAPER_FRONT=2 CKC_HOST_LANES=8 python tools/e2e_probe.py $wl >> $O 2>&1
CKC_PIPE_TAPER_FRONT=3 CKC_HOST_LANES=8 python tools/e2e_probe.py $wl >> $O 2>&1
CKC_PIPE_TAPER_FRONT=2 CKC_PIPE_TAPER=1 CKC_HOST_LANES=8 python tools/e2e_probe.py $wl >> $O 2>&1
done
cat $O
tools/gpu/gpu_k3.sh abb_dualarm_planning_b200/csrc/variants/libckc_pf.so
