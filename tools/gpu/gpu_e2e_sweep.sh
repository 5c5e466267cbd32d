#!/bin/bash
# end-to-end knobs of the host pipeline (ordered uploads, tapered last chunk, lanes, chunk size)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/e2e_sweep2.txt; : > $O
for wl in gd pcs; do
CKC_TIMELINE=1 CKC_ORDERED_UPLOAD=0 CKC_PIPE_TAPER=0 python tools/e2e_probe.py $wl >> $O 2>&1
CKC_TIMELINE=1 CKC_ORDERED_UPLOAD=1 CKC_PIPE_TAPER=0 python tools/e2e_probe.py $wl >> $O 2>&1
CKC_TIMELINE=1 CKC_ORDERED_UPLOAD=1 CKC_PIPE_TAPER=2 python tools/e2e_probe.py $wl >> $O 2>&1
for t in 1 3; do CKC_PIPE_TAPER=$t python tools/e2e_probe.py $wl >> $O 2>&1; done
CKC_PIPE_TAPER=2 CKC_HOST_LANES=8 python tools/e2e_probe.py $wl >> $O 2>&1
CKC_PIPE_TAPER=2 CKC_PIPE_CHUNK=131072 CKC_HOST_LANES=8 python tools/e2e_probe.py $wl >> $O 2>&1
CKC_PIPE_TAPER=2 CKC_PIPE_CHUNK=131072 CKC_HOST_LANES=4 python tools/e2e_probe.py $wl >> $O 2>&1
done
cat $O
