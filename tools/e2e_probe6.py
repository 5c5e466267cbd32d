"""probe: GD / PCS end-to-end (compact API) across host lane counts and pipeline chunk sizes"""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for wl in ("gd", "pcs"):
        for pc, hl in ((1 << 17, 4), (1 << 17, 8), (1 << 16, 8), (1 << 16, 16), (1 << 15, 16), (1 << 18, 4)):
            subprocess.run([sys.executable, __file__, wl, str(pc), str(hl), "0"])
    subprocess.run([sys.executable, __file__, "gd", str(1 << 17), "8", "1"])
    sys.exit(0)
wl, pc, hl, tl = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
os.environ["CKC_PIPE_CHUNK"] = str(pc); os.environ["CKC_HOST_LANES"] = str(hl)
import numpy as np, torch, ctypes
from abb_dualarm_planning_b200 import Checker, sampling
n = 1_000_000
if wl == "gd":
    q = sampling.sample_gd(1, 0, n); ik = None
else:
    q, ik = sampling.sample_pcs(1, 0, n)
hq = torch.from_numpy(q).pin_memory(); hv = torch.empty(n, dtype=torch.uint8).pin_memory()
hik = torch.from_numpy(ik).pin_memory() if ik is not None else None
cap = n // 10
hidx = torch.empty(cap, dtype=torch.int32).pin_memory(); hqv = torch.empty((cap, 12), dtype=torch.float64).pin_memory()
nv = ctypes.c_int64(0)
ck = Checker(1)
def call():
    if wl == "gd": ck._chk(ck.L.ckc_gd_validate_compact(ck.ctx, n, hq.data_ptr(), hv.data_ptr(), ctypes.byref(nv), hidx.data_ptr(), hqv.data_ptr(), cap))
    else: ck._chk(ck.L.ckc_pcs_validate_compact(ck.ctx, n, hq.data_ptr(), None, hik.data_ptr(), hv.data_ptr(), ctypes.byref(nv), hidx.data_ptr(), hqv.data_ptr(), cap))
for _ in range(3): call()
t0 = time.perf_counter()
for _ in range(8): call()
t = (time.perf_counter() - t0) / 8
print("%s pipe_chunk %7d lanes %2d: %.3f ms/step -> %.3e configs/s (valid %d)" % (wl, pc, hl, 1e3 * t, n / t, nv.value), flush=True)
if tl:
    ck.set_stage_timing(True)
    call()
    ck.L.ckc_debug_timeline(ck.ctx)
ck.close()
