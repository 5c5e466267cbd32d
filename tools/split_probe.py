"""probe: one GPU driven through a multi-device context that lists the SAME device k times (k sub-contexts, k host threads, k independent
frontiers / pipelines on one GPU) against the plain single context: 1e5-edge RBS call, 1e6-sample GD / PCS compact calls"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from abb_dualarm_planning_b200 import Checker, sampling
base = Checker(env=1, device=0)
n_e = 100000
qa, qb, iks = bench.make_rbs_edges(base, sampling, n_e, 7001)
qa, qb, iks = qa[:n_e], qb[:n_e], iks[:n_e]
h_qa = torch.from_numpy(np.ascontiguousarray(qa)).pin_memory(); h_qb = torch.from_numpy(np.ascontiguousarray(qb)).pin_memory(); h_ik = torch.from_numpy(np.ascontiguousarray(iks)).pin_memory()
n = 1_000_000
gq = torch.from_numpy(sampling.sample_gd(11, 0, n)).pin_memory()
pq_, pik_ = sampling.sample_pcs(11, 0, n); pq = torch.from_numpy(pq_).pin_memory(); pik = torch.from_numpy(pik_).pin_memory()
cap = n // 8
ref = {}
for k in (1, 2, 3, 4):
    ck = base if k == 1 else Checker(env=1, devices=[0] * k)
    h_ok = torch.empty(n_e, dtype=torch.uint8).pin_memory(); h_nc = torch.empty(n_e, dtype=torch.int32).pin_memory()
    hv = torch.empty(n, dtype=torch.uint8).pin_memory(); hidx = torch.empty(cap, dtype=torch.int32).pin_memory(); hqv = torch.empty((cap, 12), dtype=torch.float64).pin_memory(); nv = C.c_int64()
    calls = {"rbs": lambda: ck._chk(ck.L.ckc_pcs_check_motion_rbs(ck.ctx, n_e, h_qa.data_ptr(), h_qb.data_ptr(), None, h_ik.data_ptr(), h_ok.data_ptr(), h_nc.data_ptr())),
             "gd": lambda: ck._chk(ck.L.ckc_gd_validate_compact(ck.ctx, n, gq.data_ptr(), hv.data_ptr(), C.byref(nv), hidx.data_ptr(), hqv.data_ptr(), cap)),
             "pcs": lambda: ck._chk(ck.L.ckc_pcs_validate_compact(ck.ctx, n, pq.data_ptr(), None, pik.data_ptr(), hv.data_ptr(), C.byref(nv), hidx.data_ptr(), hqv.data_ptr(), cap))}
    for name, fn in calls.items():
        for _ in range(3): fn()
        R = 10 if name == "rbs" else 40
        t0 = time.perf_counter()
        for _ in range(R): fn()
        t = (time.perf_counter() - t0) / R
        sig = (int(h_ok.sum()), int(h_nc.sum())) if name == "rbs" else (int(nv.value), int(hv.sum()))
        ref.setdefault(name, sig)
        print("sub-contexts %d  %-3s %.3f ms per call  (%.3e /s)  result %s %s" % (k, name, 1e3 * t, (n_e if name == "rbs" else n) / t, sig, "same" if sig == ref[name] else "DIFFERENT"), flush=True)
    if k > 1: ck.close()
base.close()
