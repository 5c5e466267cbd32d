import sys, time, numpy as np
sys.path.insert(0, '.')
from abb_dualarm_planning_b200 import Checker, sampling
ck = Checker(1)
q, ik = sampling.sample_pcs(1, 0, 4096)
ok, valid, qo = ck.pcs_validate(q, 0, ik)
f = np.nonzero(ok)[0][:50]
for n in (1, 16, 256, 4096):
    t0 = time.perf_counter()
    for r in range(200): ck.pcs_validate(q[:n], 0, ik[:n])
    t = (time.perf_counter() - t0) / 200
    print("pcs_validate n=%5d: %.1f us per call, %.2f us per sample" % (n, 1e6 * t, 1e6 * t / n))
qa = qo[f[:1]]; 
t0 = time.perf_counter()
for r in range(200): ck.collide(qa)
print("collide n=1: %.1f us" % (1e6 * (time.perf_counter() - t0) / 200))
