"""Round-2 golden vectors produced by the GENUINE reference code (the prebuilt tests/trbspcs driven through oracle/refshim):
  tests/golden/reconstruct_rbs.npz  reconstructRBS(qa1, qa2, qb1, qb2, ac, ik, M, 0, 1, 1) on 60 edges of the rbs.npz set: bool + matrix
  tests/golden/ikproject5.npz       IKproject(q1, q2, ik, active_chain, ik_nn) on 2000 S-PCS samples, both active chains
Run in the build container (needs oracle/_ref, i.e. /root/reference mounted):  python tools/gen_golden_r2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import refshim
from abb_dualarm_planning_b200 import sampling

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
g = np.load(os.path.join(GOLD, "rbs.npz"))
sel = np.concatenate([np.nonzero(g["ok"] == 1)[0][:45], np.nonzero(g["ok"] == 0)[0][:15]])
ok, mats = refshim.reconstruct_rbs(g["qa"][sel], g["qb"][sel], 0, g["ik"][sel])
np.savez_compressed(os.path.join(GOLD, "reconstruct_rbs.npz"), qa=g["qa"][sel], qb=g["qb"][sel], ik=g["ik"][sel], ok=ok,
                    nrows=np.array([len(m) for m in mats], np.int32), rows=np.concatenate(mats))
print("reconstruct_rbs: %d edges, %d valid, rows per valid edge %.1f" % (len(sel), ok.sum(), np.mean([len(m) for m, o in zip(mats, ok) if o])))

q, ik = sampling.sample_pcs(77, 0, 60000)
okp, qp, _ = refshim.pcs_project(q, 0, ik)
base = qp[okp == 1][:1000]                                  # closed configurations ...
kk = ik[okp == 1][:1000]
rng = np.random.default_rng(3)
pert = base + rng.normal(size=base.shape) * 0.05            # ... perturbed like a planner's extension step
qin = np.concatenate([pert, pert]); ac = np.concatenate([np.zeros(1000, np.int8), np.ones(1000, np.int8)])
nn = np.stack([np.concatenate([kk, kk]), rng.integers(0, 8, 2000)], axis=1).astype(np.int8)
v, aco, iko, qo = refshim.ikproject5(qin, ac, nn)
np.savez_compressed(os.path.join(GOLD, "ikproject5.npz"), q=qin, ac=ac, ik_nn=nn, valid=v, ac_out=aco, ik=iko, q_out=qo)
print("ikproject5: %d samples, %d valid, %d switched the active chain" % (len(qin), v.sum(), (aco != ac).sum()))
