#!/usr/bin/env python
"""Generate tests/golden/* from the GENUINE reference code (oracle/_ref: the reference's prebuilt tests/trbspcs driven
by oracle/refshim, and the reference's fixture files under /root/reference).  Run in the build container only
(needs /root/reference); the resulting small fixtures are committed so that the GPU box needs nothing else.

    python tools/gen_golden.py            # everything (about 2 minutes)
"""
import json
import os
import sys
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import lib, refshim  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
REFROOT = "/root/reference"


def rbs_edges(O, seed, n_edges):
    """S-RBS protocol (tests/test_rbs_pcs.cpp:39-47): A = valid S-PCS sample, B = A + s (U - A), s in [0.01,1],
    projected on A's (ac=0, ik) and valid."""
    qa, qb, iks = [], [], []
    i0 = 0
    kin_lo = np.array([O.kin.lim[2 * j] for j in range(6)]); kin_hi = np.array([O.kin.lim[2 * j + 1] for j in range(6)])
    kin_lo[5], kin_hi[5] = -3.1416, 3.1416
    while len(qa) < n_edges:
        q, ik = O.sample_pcs(seed, i0, 20000)
        ok, valid, qo = O.pcs_validate(q, 0, ik)
        for idx in np.nonzero(valid)[0]:
            gi = i0 + int(idx)
            a = qo[idx]
            u = np.array([O.L.ckco_uniform(seed + 1000, gi, k) for k in range(6)])
            s = 0.01 + 0.99 * O.L.ckco_uniform(seed + 1000, gi, 6)
            U = kin_lo + u * (kin_hi - kin_lo)
            b = a.copy(); b[:6] = a[:6] + s * (U - a[:6])
            okb, vb, qbo = O.pcs_validate(b[None], 0, ik[idx:idx + 1])
            if vb[0]:
                qa.append(a); qb.append(qbo[0]); iks.append(ik[idx])
                if len(qa) == n_edges:
                    break
        i0 += 20000
    return np.array(qa), np.array(qb), np.array(iks, np.int8)



def main():
    assert refshim.available(), "run `make -C oracle ref` first (needs /root/reference)"
    os.makedirs(GOLD, exist_ok=True)
    O = lib.Oracle(1)
    rng = np.random.default_rng(20261019)

    # 1. the 1e6-sample S-PCS seed set of the north star: masks from the genuine reference
    N = 1_000_000
    ok, valid, secs = refshim.spcs_seeded(1, 0, N)
    np.savez_compressed(os.path.join(GOLD, "spcs_seed1_1e6_mask.npz"), seed=1, n=N, ok_bits=np.packbits(ok), valid_bits=np.packbits(valid))
    print("S-PCS 1e6: feasible %d valid %d ref %.1f s (%.1f us/sample)" % (ok.sum(), valid.sum(), secs, 1e6 * secs / N))

    # work-model counters from the genuine PQP on the feasible configs of the first 250k samples
    q, ik = O.sample_pcs(1, 0, 250000)
    okp, qo, _ = refshim.pcs_project(q, 0, ik)
    assert (okp == ok[:250000]).all()
    qf = qo[okp == 1]
    flags, cnt, refc = refshim.pair_counters(qf)
    wm = dict(seed=1, n_samples=250000, n_feasible=int(len(qf)), f_ik=float(len(qf) / 250000.0), n_colliding=int((refc == 1).sum()),
              n_bv_mean=float(cnt[:, 0].mean()), n_tri_mean=float(cnt[:, 1].mean()),
              n_bv_free=float(cnt[refc == 0, 0].mean()), n_tri_free=float(cnt[refc == 0, 1].mean()),
              n_bv_coll=float(cnt[refc == 1, 0].mean()), n_tri_coll=float(cnt[refc == 1, 1].mean()),
              n_bv_max=int(cnt[:, 0].max()), n_tri_max=int(cnt[:, 1].max()),
              source="PQP v1.3 counters (num_bv_tests/num_tri_tests) read from the reference's prebuilt tests/trbspcs, root tests excluded",
              ref_seconds_per_sample=secs / N)
    json.dump(wm, open(os.path.join(GOLD, "work_model.json"), "w"), indent=1)
    print("work model", wm)

    # 2. projection golden (both active chains), joint values from the binary
    q, ik = O.sample_pcs(3, 0, 60000)
    ok0, q0, _ = refshim.pcs_project(q, 0, ik)
    ok1, q1, _ = refshim.pcs_project(q, 1, ik)
    keep = np.unique(np.concatenate([np.nonzero(ok0)[0], np.nonzero(ok1)[0], np.arange(2000)]))
    np.savez_compressed(os.path.join(GOLD, "pcs_project.npz"), q=q[keep], ik=ik[keep], ok_ac0=ok0[keep], q_ac0=q0[keep], ok_ac1=ok1[keep], q_ac1=q1[keep])
    print("projection golden: %d rows, %d / %d feasible" % (len(keep), ok0.sum(), ok1.sum()))

    # 3. collision golden: feasible configs, per-pair PQP flags + binary collision_state
    qf3 = q0[ok0 == 1][:1500]
    flags, cnt, refc = refshim.pair_counters(qf3)
    np.savez_compressed(os.path.join(GOLD, "collide.npz"), q=qf3, pair_flags=np.packbits(flags, axis=1), counters=cnt, collision_state=refc)
    print("collision golden: %d configs, %d colliding" % (len(qf3), refc.sum()))

    # 4. TriDist golden (incl. degenerate and touching triangles)
    n = 4000
    S = rng.normal(size=(n, 3, 3)) * 10
    T = rng.normal(size=(n, 3, 3)) * 10 + rng.normal(size=(n, 1, 3)) * rng.uniform(0, 30, size=(n, 1, 1))
    S[:1000, 1] = S[:1000, 0] + (S[:1000, 2] - S[:1000, 0]) * 0.5
    T[1000:2000, 0] = S[1000:2000, 0]
    d, pq = refshim.tridist(S, T)
    np.savez_compressed(os.path.join(GOLD, "tridist.npz"), S=S, T=T, d=d)

    # 5. RBS golden
    qa, qb, iks = rbs_edges(O, 5, 300)
    okr, nodes, secs = refshim.rbs(qa, qb, 0, iks)
    np.savez_compressed(os.path.join(GOLD, "rbs.npz"), qa=qa, qb=qb, ik=iks, ok=okr, nchecks=nodes)
    print("RBS golden: %d edges, %.1f%% succeed, %.1f checks/edge, ref %.2f ms/edge" % (len(qa), 100 * okr.mean(), nodes.mean(), 1e3 * secs / len(qa)))

    # 6. identify_state_ik golden: closed configs (both arms consistent) and perturbed ones
    qc = q0[ok0 == 1][:600].copy()
    qc[300:] += rng.normal(size=qc[300:].shape) * 0.03
    ikid = refshim.identify_ik(qc)
    np.savez_compressed(os.path.join(GOLD, "identify_ik.npz"), q=qc, ik=ikid)

    # 6c. isValid(const ob::State*) (PCS.cpp:323-357) through the binary on closed, perturbed and valid configurations
    gi, gr = np.load(os.path.join(GOLD, "identify_ik.npz")), np.load(os.path.join(GOLD, "rbs.npz"))
    r3 = np.random.default_rng(3)
    qv = np.concatenate([gi["q"], gr["qa"][:200], gr["qa"][:200] + r3.normal(size=(200, 12)) * 0.01, gr["qb"][:200] + r3.normal(size=(200, 12)) * 0.04])
    vfull = refshim.isvalid_full(qv)
    idv = refshim.identify_ik(qv)
    np.savez_compressed(os.path.join(GOLD, "isvalid_full.npz"), q=qv, valid=vfull, ik=idv)
    print("isValid(state): %d configs, %d identified, %d valid" % (len(qv), (idv >= 0).all(1).sum(), vfull.sum()))

    # 6b. env 2 (cones, D = 1200, L = 500, "T7z > 800" rule): masks of 60000 S-PCS samples from the genuine binary
    O2 = lib.Oracle(2)
    q2, ik2 = O2.sample_pcs(4, 0, 60000)
    ok2, valid2, _, _ = refshim.pcs_validate(q2, 0, ik2, env=2)
    np.savez_compressed(os.path.join(GOLD, "spcs_env2_seed4_60k_mask.npz"), seed=4, n=60000, ok_bits=np.packbits(ok2), valid_bits=np.packbits(valid2))
    print("env 2: feasible %d valid %d" % (ok2.sum(), valid2.sum()))

    # 7. reference fixture files (data): solved path, RX accepted configurations
    path = np.loadtxt(os.path.join(REFROOT, "paths", "path.txt"), skiprows=1)
    cs, _ = refshim.collide(path)
    np.savez_compressed(os.path.join(GOLD, "ref_path.npz"), path=path, collision_state=cs)
    print("paths/path.txt: %d rows, %d colliding" % (len(path), cs.sum()))
    for eps, env, name, out in ((0.7, 1, "rlx_rand_confs_eps0.700000.txt", "rx_confs_eps0.7.npz"),
                                (1.0, 1, "rlx_rand_confs_eps1.000000.txt", "rx_confs_eps1.0.npz"),
                                (0.5, 2, "rlx_rand_confs_eps0.5_env2.txt", "rx_confs_eps0.5_env2.npz")):
        a = np.loadtxt(os.path.join(REFROOT, "tests", "results", "data", name))[:, :12]
        if len(a) > 2000:
            a = a[:: len(a) // 2000][:2000]
        np.savez_compressed(os.path.join(GOLD, out), q=a, eps=eps, env=env)
        print(name, a.shape)


if __name__ == "__main__":
    main()
