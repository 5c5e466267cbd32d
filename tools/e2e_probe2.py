import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from abb_dualarm_planning_b200 import Checker, sampling
n = 1_000_000
q = sampling.sample_gd(1, 0, n)
hq = torch.from_numpy(q).pin_memory(); ho = torch.empty((n, 12), dtype=torch.float64).pin_memory(); hv = torch.empty(n, dtype=torch.uint8).pin_memory()
for pc in (1 << 17, 1 << 18, 1 << 20):
    os.environ["CKC_PIPE_CHUNK"] = str(pc)
    ck = Checker(1)
    def call(qo=True): ck._chk(ck.L.ckc_gd_validate(ck.ctx, n, hq.data_ptr(), ho.data_ptr() if qo else None, hv.data_ptr()))
    call(); call()
    ck.set_stage_timing(True)
    t0 = time.perf_counter(); call(); t = time.perf_counter() - t0
    st = ck.stage_times(); ck.set_stage_timing(False)
    print("chunk %7d: %.2f ms; stage sums:" % (pc, 1e3 * t), {k: (round(v["ms"], 2), v["launches"]) for k, v in st.items()})
    t0 = time.perf_counter(); call(False); t = time.perf_counter() - t0
    print("     without q_out D2H: %.2f ms" % (1e3 * t))
    ck.close()
