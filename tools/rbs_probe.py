import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abb_dualarm_planning_b200 import Checker, sampling
import bench
ck = Checker(1)
for ne in (2000, 20000, 100000):
    qa, qb, iks = bench.make_rbs_edges(ck, sampling, ne, 4242)
    ck.pcs_check_motion_rbs(qa[:500], qb[:500], 0, iks[:500])
    ck.reset_stats(); ck.set_stage_timing(True)
    t0 = time.perf_counter(); ok, nc = ck.pcs_check_motion_rbs(qa, qb, 0, iks); t = time.perf_counter() - t0
    st = ck.stage_times(); ck.set_stage_timing(False); s = ck.stats()
    print("%6d edges: %.1f ms -> %.3e edges/s; ok %.1f%% checks/edge %.1f; stages ms %s; launches %d" % (ne, 1e3 * t, ne / t, 100 * ok.mean(), nc.mean(), {k: round(v["ms"], 2) for k, v in st.items()}, s["kernel_launches"]))
