import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from abb_dualarm_planning_b200 import Checker, sampling
n = 1_000_000
for wl in ("gd", "pcs"):
    if wl == "gd": q = sampling.sample_gd(1, 0, n); ik = None
    else: q, ik = sampling.sample_pcs(1, 0, n)
    hq = torch.from_numpy(q).pin_memory(); ho = torch.empty((n, 12), dtype=torch.float64).pin_memory(); hv = torch.empty(n, dtype=torch.uint8).pin_memory(); hk = torch.empty(n, dtype=torch.uint8).pin_memory()
    hik = torch.from_numpy(ik).pin_memory() if ik is not None else None
    ref = None
    for zc in ("0", "1"):
        os.environ["CKC_ZERO_COPY"] = zc
        ck = Checker(1)
        def call():
            if wl == "gd": ck._chk(ck.L.ckc_gd_validate(ck.ctx, n, hq.data_ptr(), ho.data_ptr(), hv.data_ptr()))
            else: ck._chk(ck.L.ckc_pcs_validate(ck.ctx, n, hq.data_ptr(), None, hik.data_ptr(), ho.data_ptr(), hk.data_ptr(), hv.data_ptr()))
        call(); call()
        t0 = time.perf_counter()
        for _ in range(5): call()
        t = (time.perf_counter() - t0) / 5
        v = hv.numpy().copy(); o = ho.numpy().copy()
        if ref is None: ref = (v, o)
        print("%s zero_copy=%s: %.2f ms/step -> %.3e configs/s  valid=%d  same_as_staged=%s" % (wl, zc, 1e3 * t, n / t, v.sum(), (v == ref[0]).all() and (o[v == 1] == ref[1][v == 1]).all()))
        ck.close()
