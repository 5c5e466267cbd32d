"""probe: end-to-end rate of the host-buffer compact entry points (pinned host buffers, 1e6 samples per call) under the current
environment knobs (CKC_PIPE_TAPER, CKC_HOST_LANES, CKC_PIPE_CHUNK): tools/e2e_probe.py [gd|pcs] [n]"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from abb_dualarm_planning_b200 import Checker, sampling
wl = sys.argv[1] if len(sys.argv) > 1 else "gd"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
ck = Checker(env=1, device=0)
NB = 3
hq, hik = [], []
for b in range(NB):
    if wl == "gd": q = sampling.sample_gd(11 + b, 0, n); ik = None
    else: q, ik = sampling.sample_pcs(11 + b, 0, n)
    hq.append(torch.from_numpy(q).pin_memory()); hik.append(torch.from_numpy(ik).pin_memory() if ik is not None else None)
cap = n // 8
hv = torch.empty(n, dtype=torch.uint8).pin_memory(); hidx = torch.empty(cap, dtype=torch.int32).pin_memory(); hqv = torch.empty((cap, 12), dtype=torch.float64).pin_memory()
nv = C.c_int64()
def call(b):
    if wl == "gd": ck._chk(ck.L.ckc_gd_validate_compact(ck.ctx, n, hq[b].data_ptr(), hv.data_ptr(), C.byref(nv), hidx.data_ptr(), hqv.data_ptr(), cap))
    else: ck._chk(ck.L.ckc_pcs_validate_compact(ck.ctx, n, hq[b].data_ptr(), None, hik[b].data_ptr(), hv.data_ptr(), C.byref(nv), hidx.data_ptr(), hqv.data_ptr(), cap))
for i in range(5): call(i % NB)
best = 1e9; tot = 0.0; R = 60
for i in range(R):
    t0 = time.perf_counter(); call(i % NB); dt = time.perf_counter() - t0
    best = min(best, dt); tot += dt
knobs = " ".join("%s=%s" % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith("CKC_"))
print("%-3s n=%d  mean %.3f ms (%.3e /s)  best %.3f ms  valid %d  sum(mask) %d  | %s" % (wl, n, 1e3 * tot / R, n * R / tot, 1e3 * best, nv.value, int(hv.sum()), knobs), flush=True)
if os.environ.get("CKC_TIMELINE"):
    ck.L.ckc_set_stage_timing(ck.ctx, 1); call(0); sys.stderr.flush(); ck.L.ckc_debug_timeline(ck.ctx); ck.L.ckc_set_stage_timing(ck.ctx, 0)
ck.close()
