"""probe: PCIe copy bandwidth and the host-API pipeline at several chunk sizes (run on the GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
n = 1_000_000
h = torch.empty((n, 12), dtype=torch.float64).pin_memory(); d = torch.empty((n, 12), dtype=torch.float64, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 5
    print("%s 96 MB: %.2f ms = %.1f GB/s" % (name, 1e3 * t, 0.096 / t))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty((n, 12), dtype=torch.float64).pin_memory(); d2 = torch.empty((n, 12), dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 5
print("H2D || D2H 96 MB each: %.2f ms" % (1e3 * t))
from abb_dualarm_planning_b200 import Checker, sampling
for wl in ("gd", "pcs"):
    for pc in (1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 20):
        os.environ["CKC_PIPE_CHUNK"] = str(pc)
        ck = Checker(1)
        if wl == "gd":
            q = sampling.sample_gd(1, 0, n); ik = None
        else:
            q, ik = sampling.sample_pcs(1, 0, n)
        hq = torch.from_numpy(q).pin_memory(); ho = torch.empty((n, 12), dtype=torch.float64).pin_memory(); hv = torch.empty(n, dtype=torch.uint8).pin_memory(); hk = torch.empty(n, dtype=torch.uint8).pin_memory()
        hik = torch.from_numpy(ik).pin_memory() if ik is not None else None
        def call():
            if wl == "gd": ck._chk(ck.L.ckc_gd_validate(ck.ctx, n, hq.data_ptr(), ho.data_ptr(), hv.data_ptr()))
            else: ck._chk(ck.L.ckc_pcs_validate(ck.ctx, n, hq.data_ptr(), None, hik.data_ptr(), ho.data_ptr(), hk.data_ptr(), hv.data_ptr()))
        call(); call()
        t0 = time.perf_counter()
        for _ in range(5): call()
        t = (time.perf_counter() - t0) / 5
        print("%s pipe_chunk %7d: %.2f ms/step -> %.3e configs/s" % (wl, pc, 1e3 * t, n / t))
        ck.close()
