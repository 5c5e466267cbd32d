import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from abb_dualarm_planning_b200 import Checker, sampling
n = 1_000_000
q = sampling.sample_gd(1, 0, n)
hq = torch.from_numpy(q).pin_memory(); ho = torch.empty((n, 12), dtype=torch.float64).pin_memory(); hv = torch.empty(n, dtype=torch.uint8).pin_memory()
for pc in (1 << 18,):
    os.environ["CKC_PIPE_CHUNK"] = str(pc)
    ck = Checker(1)
    def call(): ck._chk(ck.L.ckc_gd_validate(ck.ctx, n, hq.data_ptr(), ho.data_ptr(), hv.data_ptr()))
    call(); call()
    ck.set_stage_timing(True)
    t0 = time.perf_counter(); call(); t = time.perf_counter() - t0
    print("chunk %d: %.2f ms" % (pc, 1e3 * t), flush=True)
    ck.L.ckc_debug_timeline(ck.ctx)
    ck.close()
