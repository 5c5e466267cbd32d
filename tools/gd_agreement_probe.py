"""probe: how closely the device GD projection agrees with the CPU restatement (fractions behind the tolerances of tests/test_gpu_parity.py)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abb_dualarm_planning_b200 import Checker, sampling
from oracle import lib as olib
for env, n in ((1, 100000), (2, 100000)):
    ck = Checker(env=env); orc = olib.Oracle(env)
    q = sampling.sample_gd(1, 0, n)
    ok_o, conv_o, it_o, q_o = orc.gd_project(q)
    ok, conv, it, qo = ck.gd_project(q)
    d = np.abs(qo - q_o); d = np.minimum(d, np.abs(d - 2 * 3.1416)); dm = d.max(axis=1)
    print("env %d n %d: conv all %s | ok mismatches %d | iteration: equal %.6f within one %.6f | joints: <1e-9 %.6f <1e-6 %.6f <1e-4 %.6f max %.3g" % (
        env, n, bool(conv.all() and conv_o.all()), int((ok != ok_o).sum()), (it == it_o).mean(), (np.abs(it - it_o) <= 1).mean(),
        (dm < 1e-9).mean(), (dm < 1e-6).mean(), (dm < 1e-4).mean(), dm.max()), flush=True)
    v, _ = ck.gd_validate(q); vo, _ = orc.gd_validate(q)
    print("   validate mask mismatches %d of %d (valid %d)" % (int((v != vo).sum()), n, int(v.sum())), flush=True)
    ck.close()
