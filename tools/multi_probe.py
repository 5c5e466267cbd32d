"""probe: the in-library multi-device context (ckc_create_multi) driven from ONE process: S-PCS / S-GD host-buffer calls over ndev GPUs
(one host thread per device inside the library) against one device; masks must be identical.  usage: multi_probe.py [ndev]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from abb_dualarm_planning_b200 import Checker, sampling
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
n = 4_000_000
q, ik = sampling.sample_pcs(5, 0, n)
hq = torch.from_numpy(q).pin_memory(); hik = torch.from_numpy(ik).pin_memory()
hv = torch.empty(n, dtype=torch.uint8).pin_memory(); hok = torch.empty(n, dtype=torch.uint8).pin_memory(); hqo = torch.empty((n, 12), dtype=torch.float64).pin_memory()
res = {}
for nd in sorted(set([1, 2, ndev])):
    if nd > torch.cuda.device_count(): continue
    ck = Checker(env=1, devices=list(range(nd)))
    def call():
        ck._chk(ck.L.ckc_pcs_validate(ck.ctx, n, hq.data_ptr(), None, hik.data_ptr(), hqo.data_ptr(), hok.data_ptr(), hv.data_ptr()))
    call(); call()
    t0 = time.perf_counter()
    for _ in range(5): call()
    t = (time.perf_counter() - t0) / 5
    res[nd] = hv.numpy().copy()
    print("ckc_create_multi ndev=%d: ckc_pcs_validate %d samples %.2f ms  -> %.3e configs/s end to end (full output), valid %d" % (nd, n, 1e3 * t, n / t, int(res[nd].sum())), flush=True)
    ck.close()
ks = sorted(res)
print("masks identical across device counts:", all((res[k] == res[ks[0]]).all() for k in ks))
