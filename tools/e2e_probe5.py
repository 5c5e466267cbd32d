"""probe: GD end-to-end (compact API) across pipeline chunk sizes / samples-per-lane targets, plus one stage timeline"""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for pc in (1 << 16, 1 << 17, 1 << 18):
        for spl in (1, 2, 3, 6):
            subprocess.run([sys.executable, __file__, str(pc), str(spl), "0"])
    subprocess.run([sys.executable, __file__, str(1 << 17), "6", "1"])
    subprocess.run([sys.executable, __file__, str(1 << 17), "2", "1"])
    sys.exit(0)
pc, spl, tl = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
os.environ["CKC_PIPE_CHUNK"] = str(pc); os.environ["CKC_GD_SPL"] = str(spl)
import numpy as np, torch
from abb_dualarm_planning_b200 import Checker, sampling
n = 1_000_000
q = sampling.sample_gd(1, 0, n)
hq = torch.from_numpy(q).pin_memory(); hv = torch.empty(n, dtype=torch.uint8).pin_memory()
cap = n // 10
hidx = torch.empty(cap, dtype=torch.int32).pin_memory(); hqv = torch.empty((cap, 12), dtype=torch.float64).pin_memory()
import ctypes
nv = ctypes.c_int64(0)
ck = Checker(1)
def call(): ck._chk(ck.L.ckc_gd_validate_compact(ck.ctx, n, hq.data_ptr(), hv.data_ptr(), ctypes.byref(nv), hidx.data_ptr(), hqv.data_ptr(), cap))
for _ in range(3): call()
t0 = time.perf_counter()
for _ in range(8): call()
t = (time.perf_counter() - t0) / 8
print("pipe_chunk %7d spl %d: %.3f ms/step -> %.3e configs/s (valid %d)" % (pc, spl, 1e3 * t, n / t, nv.value), flush=True)
if tl:
    ck.set_stage_timing(True)
    call()
    ck.L.ckc_debug_timeline(ck.ctx)
ck.close()
