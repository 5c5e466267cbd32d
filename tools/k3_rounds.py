"""probe: distribution of k_collide's work per configuration (box rounds, triangle steps) on the IK-feasible S-PCS set, and the
kernel time of the heaviest configurations alone; run on the GPU box"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abb_dualarm_planning_b200 import Checker, sampling
ck = Checker(1)
q, ik = sampling.sample_pcs(1, 0, 4_000_000)
ok, qo = ck.pcs_project(q, 0, ik)
feas = np.ascontiguousarray(qo[ok == 1])
n = len(feas)
r = np.zeros(n, np.uint32)
ck.L.ckc_debug_collide_rounds.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
rc = ck.L.ckc_debug_collide_rounds(ck.ctx, n, feas.ctypes.data, r.ctypes.data)
box, tri = r & 0xffff, r >> 16
hit = ck.collide(feas)
print("rc", rc, "n", n, "box rounds mean %.2f p50 %d p90 %d p99 %d p99.9 %d max %d" % (box.mean(), *np.percentile(box, [50, 90, 99, 99.9]), box.max()))
print("tri steps mean %.2f p50 %d p90 %d p99 %d p99.9 %d max %d" % (tri.mean(), *np.percentile(tri, [50, 90, 99, 99.9]), tri.max()))
cost = box + 2 * tri
order = np.argsort(-cost.astype(np.int64))
print("heaviest 15 (box, tri, hit):", [(int(box[i]), int(tri[i]), int(hit[i])) for i in order[:15]])
for label, sel in (("heaviest 64", order[:64]), ("heaviest 1000", order[:1000]), ("lightest 30000", order[-30000:]), ("random 30000", np.random.default_rng(0).permutation(n)[:30000]),
                   ("random 30000 minus heaviest 300", np.setdiff1d(np.random.default_rng(0).permutation(n)[:30000], order[:300]))):
    arr = np.ascontiguousarray(feas[sel])
    for _ in range(2): ck.collide(arr)
    ck.set_stage_timing(True)
    for _ in range(5): ck.collide(arr)
    st = ck.stage_times(); ck.set_stage_timing(False)
    print("%-34s n=%6d collide %.3f ms" % (label, len(arr), st["collide"]["ms"] / 5))
ck.close()
