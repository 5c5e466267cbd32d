"""probe: K3 (k_collide) time and work counters on IK-feasible S-PCS configurations and on RBS-like (mostly free) ones,
for several values of an environment knob given as NAME=v1,v2,...   usage: k3_probe.py CKC_WIDE_CNT=0,8,16,32"""
import os, sys, subprocess, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 2 and "=" in sys.argv[1]:
    name, vals = sys.argv[1].split("=")
    for v in vals.split(","):
        subprocess.run([sys.executable, __file__, name, v])
    sys.exit(0)
name, v = sys.argv[1], sys.argv[2]
os.environ[name] = v
import numpy as np
from abb_dualarm_planning_b200 import Checker, sampling
import time
_t0 = time.time(); ck = Checker(1); print('%s=%s create %.2f s' % (name, v, time.time() - _t0))
q, ik = sampling.sample_pcs(1, 0, 4_000_000)
ok, qo = ck.pcs_project(q, 0, ik)
feas = qo[ok == 1]                                   # ~95k IK-feasible configurations, half of them colliding
hit = ck.collide(feas)
free = feas[hit == 0]
for label, arr in (("feasible", feas), ("free", free), ("feas31k", feas[:31500])):
    for _ in range(2): ck.collide(arr)
    ck.reset_stats(); ck.set_stage_timing(True)
    for _ in range(5): h = ck.collide(arr)
    st = ck.stage_times(); s = ck.stats(); ck.set_stage_timing(False)
    ms = st["collide"]["ms"] / 5
    print("%s=%s %-8s n=%d hit=%d  collide %.3f ms  (%.1f ns/config)  bv/config %.1f tri/config %.2f  ovf %d  md5 %s" % (
        name, v, label, len(arr), int(h.sum()), ms, 1e6 * ms / len(arr), s["bv_tests"] / s["collision_calls"], s["tri_tests"] / s["collision_calls"],
        s["queue_overflows"], hashlib.md5(h.tobytes()).hexdigest()[:8]), flush=True)
ck.close()
