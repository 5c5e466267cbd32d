"""GPU: the C++ drop-in StateValidityChecker classes (abb_dualarm_planning_b200/host) driven the way the reference's own
tests/test_rbs_pcs.cpp / test_rbs_gd.cpp drive the original classes, every answer compared with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu
HOST = os.path.join(REPO, "abb_dualarm_planning_b200", "host")
PACK = os.path.join(REPO, "abb_dualarm_planning_b200", "data", "meshes.ckcmesh")


def _run(binary, inp, out, **extra_env):
    env = dict(os.environ, CKC_MESH_PATH=PACK, **extra_env)
    r = subprocess.run([os.path.join(HOST, binary), inp, out], env=env, capture_output=True, text=True, timeout=600, cwd=os.path.dirname(out))
    assert r.returncode == 0, r.stderr[-2000:]


def test_pcs_dropin_class(tmp_path, oracle, gold):
    from abb_dualarm_planning_b200 import sampling
    g = gold("rbs.npz")
    n = 60
    q, ik = sampling.sample_pcs(31, 0, 400)
    rows = []
    for i in range(n):                      # valid RBS endpoints from the golden set + random (mostly infeasible) samples
        rows.append(np.concatenate([g["qa"][i], [g["ik"][i]], g["qb"][i]]))
    for i in range(n):
        rows.append(np.concatenate([q[i], [ik[i]], q[i + 100]]))
    rows = np.array(rows)
    inp, out = str(tmp_path / "in.txt"), str(tmp_path / "out.txt")
    os.makedirs(tmp_path / "paths", exist_ok=True)
    with open(inp, "w") as f:
        f.write("%d\n" % len(rows))
        for r in rows:
            f.write(" ".join(repr(float(x)) for x in r[:12]) + " %d " % int(r[12]) + " ".join(repr(float(x)) for x in r[13:]) + "\n")
    _run("test_dropin_pcs", inp, out)
    lines = open(out).read().strip().split("\n")
    # the same driver on a multi-device context (CKC_DEVICES: the drop-in class then owns one sub-context per listed device and the
    # library splits every batch over them; with one GPU in the box the same device twice): identical answers
    import torch
    out2 = str(tmp_path / "out_multi.txt")
    _run("test_dropin_pcs", inp, out2, CKC_DEVICES="0,1" if torch.cuda.device_count() >= 2 else "0,0")
    assert open(out2).read().strip().split("\n")[:len(rows)] == lines[:len(rows)]
    res = np.array([[float(x) for x in ln.split()] for ln in lines[:len(rows)]])
    qq, kk, qb = rows[:, :12], rows[:, 12].astype(np.int8), rows[:, 13:]
    ok_o, q_o = oracle.pcs_project(qq, 0, kk)
    assert (res[:, 0] == ok_o).all()
    m = ok_o == 1
    assert np.abs(res[m][:, 6:12] - q_o[m][:, 6:]).max() < 1e-6
    coll_o = oracle.collision_state(q_o[m])
    assert (res[m][:, 1] == coll_o).all()
    id_o = oracle.identify_state_ik(q_o[m])
    assert (res[m][:, 2:4] == id_o).all()
    _, valid_o, _ = oracle.pcs_validate(qq, 0, kk)
    assert (res[:, 4] == valid_o).all()
    # isValid(const ob::State*) on the projected configuration and on one opened by 4 mrad in q2[1]
    full = q_o[m].copy(); opened = q_o[m].copy(); opened[:, 7] += 0.004
    assert (res[m][:, 12] == oracle.pcs_is_valid_full(full)).all() and res[m][:, 12].sum() >= n // 2
    assert (res[m][:, 13] == oracle.pcs_is_valid_full(opened)).all()
    # RBS on rows whose two endpoints are valid
    _, vb, qbo = oracle.pcs_validate(qb, 0, kk)
    both = (valid_o == 1) & (vb == 1)
    rbs_o, _ = oracle.pcs_check_motion_rbs(q_o[both], qbo[both], 0, kk[both])
    assert (res[both][:, 5] == rbs_o).all() and both.sum() >= n // 2
    assert (res[~both][:, 5] == 0).all()
    # sampler: returns a closed configuration accepted by the reference's (quirky) acceptance rule
    s = [float(x) for x in lines[len(rows)].split()[1:]]
    hit, lim, qs = int(s[0]), int(s[1]), np.array(s[2:])
    assert not (hit and not lim)
    assert (oracle.identify_state_ik(qs[None])[0] >= 0).any()
    # LogPerf2file wrote the 16-line record of StateValidityCheckerPCS.cpp:628-651
    assert len(open(tmp_path / "paths" / "perf_log.txt").read().split()) == 16


def test_gd_dropin_class(tmp_path, oracle):
    from abb_dualarm_planning_b200 import sampling
    q = sampling.sample_gd(41, 0, 300)
    inp, out = str(tmp_path / "in.txt"), str(tmp_path / "out.txt")
    with open(inp, "w") as f:
        f.write("%d\n" % len(q))
        for r in q:
            f.write(" ".join(repr(float(x)) for x in r) + "\n")
    _run("test_dropin_gd", inp, out)
    lines = open(out).read().strip().split("\n")
    res = np.array([[float(x) for x in ln.split()] for ln in lines[:len(q)]])
    ok_o, conv_o, it_o, q_o = oracle.gd_project(q)
    valid_o, _ = oracle.gd_validate(q)
    assert (res[:, 0] == ok_o).all() and (res[:, 1] == valid_o).all()
    d = np.abs(res[:, 2:] - q_o); d = np.minimum(d, np.abs(d - 2 * 3.1416))
    assert d.max() < 1e-4 and (d.max(axis=1) < 2e-5).mean() >= 0.99      # see test_gd_projection_vs_oracle for the tolerance
    batch = np.array([int(x) for x in lines[len(q)].split()[1:]])
    assert (batch == valid_o).all()


def _singular_rbs_oracle(oracle, a, b, ac, ik):
    """RSS.cpp:586-613 restated with oracle primitives: singular stop rule at the top level only"""
    if np.linalg.norm(a - b) < 0.05:
        return True
    m = (a + b) / 2
    ok, valid, mq = oracle.pcs_validate(m[None], ac, ik)
    if not valid[0]:
        return False
    if np.linalg.norm(mq[0] - a) < 15 * 0.05:
        return True
    r, _ = oracle.pcs_check_motion_rbs(np.stack([a, mq[0]]), np.stack([mq[0], b]), ac, ik)
    return bool(r[0] and r[1])


def test_rss_dropin_class(tmp_path, oracle):
    """sampleSingular + the singular-endpoint RBS variant of StateValidityCheckerRSS"""
    out = str(tmp_path / "out.txt")
    env = dict(os.environ, CKC_MESH_PATH=PACK)
    r = subprocess.run([os.path.join(HOST, "test_dropin_rss"), out], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = np.array([[float(x) for x in ln.split()] for ln in open(out).read().strip().split("\n")])
    assert len(rows) == 12
    n_rbs = 0
    for row in rows:
        ida, proj, rbs = row[0:2].astype(int), int(row[2]), int(row[3])
        a, b = row[4:16], row[16:28]
        # the singular sample: passive arm singular (q5 = 0 or the q3 = -PI_/2 modes), closed and collision free
        assert (np.abs(a[10]) < 1e-12) or (np.abs(a[8] + 3.1416 / 2) < 1e-12)
        assert (oracle.identify_state_ik(a[None])[0] == ida).all() and ida[1] >= 0
        assert oracle.collision_state(a[None])[0] == 0
        if proj:
            assert rbs == int(_singular_rbs_oracle(oracle, a, b, 1, ida[1]))
            n_rbs += 1
    assert n_rbs >= 1


def test_hb_dropin_class(tmp_path, oracle):
    """hybrid flavour (StateValidityCheckerHB): GD projection + collision, then identify_state_ik and the APC local connection"""
    from abb_dualarm_planning_b200 import sampling
    qs = sampling.sample_gd(51, 0, 40000)
    v, qp = oracle.gd_validate(qs)
    seeds = qs[v == 1][:40]                                   # seeds whose projection is valid
    rng = np.random.default_rng(7)
    qa = seeds; qb = oracle.gd_validate(seeds)[1] + rng.normal(size=seeds.shape) * 0.05      # a second seed next to the projected point
    inp, out = str(tmp_path / "in.txt"), str(tmp_path / "out.txt")
    with open(inp, "w") as f:
        f.write("%d\n" % len(qa))
        for a, b in zip(qa, qb):
            f.write(" ".join(repr(float(x)) for x in a) + " " + " ".join(repr(float(x)) for x in b) + "\n")
    _run("test_dropin_hb", inp, out)
    lines = open(out).read().strip().split("\n")
    res = np.array([[float(x) for x in ln.split()] for ln in lines[:len(qa)]])
    va_o, pa = oracle.gd_validate(qa); vb_o, pb = oracle.gd_validate(qb)
    assert (res[:, 0] == va_o).all() and (res[:, 1] == vb_o).all() and va_o.sum() >= 30
    both = (va_o == 1) & (vb_o == 1)
    ida = oracle.identify_state_ik(pa[both]); idb = oracle.identify_state_ik(pb[both])
    assert (res[both][:, 2:4] == ida).all() and (res[both][:, 4:6] == idb).all()
    d = np.abs(res[both][:, 7:19] - pa[both]); d = np.minimum(d, np.abs(d - 2 * 3.1416))
    assert (d.max(axis=1) < 2e-5).mean() >= 0.95
    same = both.copy(); same[both] = (ida[:, 0] >= 0) & (ida[:, 0] == idb[:, 0])
    ia = oracle.identify_state_ik(pa[same])[:, 0].astype(np.int8)
    # the class ran the bisection on ITS OWN projected points (columns 7:19 and 19:31): the oracle's bisection of exactly those end points
    # must give the same answer for every edge (no tolerance: same inputs, bit-exact PCS path)
    rbs_o, _ = oracle.pcs_check_motion_rbs(res[same][:, 7:19], res[same][:, 19:31], 0, ia)
    assert (res[same][:, 6] == rbs_o).all() and same.sum() >= 10
    s = np.array([float(x) for x in lines[len(qa)].split()[1:]])
    assert oracle.gd_validate(s[None])[0][0] == 1             # the GD sampler returns a valid (closed, in limits, collision free) point


def _rx_rbs_oracle(oracle, a, b, eps, depth=0):
    """RX.cpp:216-243 restated with oracle primitives: the midpoint is NOT projected, only tested"""
    if np.linalg.norm(a - b) < 0.05:
        return True
    if depth > 150:
        return False
    m = (a + b) / 2
    ok, _ = oracle.rx_check(m[None], eps)
    if not (ok[0] and oracle.check_angle_limits(m[None])[0] and not oracle.collision_state(m[None])[0]):
        return False
    return _rx_rbs_oracle(oracle, a, m, eps, depth + 1) and _rx_rbs_oracle(oracle, m, b, eps, depth + 1)


def test_rx_dropin_class(tmp_path, oracle, gold):
    """relaxed-constraint flavour (StateValidityCheckerRX): residual test, validity, bisection over unprojected midpoints"""
    g = gold("rx_confs_eps0.7.npz")
    eps = 0.7
    qa = g["q"][:60]
    rng = np.random.default_rng(11)
    qb = qa + rng.normal(size=qa.shape) * np.where(np.arange(60)[:, None] % 2 == 0, 0.01, 0.1)      # short and longer edges
    inp, out = str(tmp_path / "in.txt"), str(tmp_path / "out.txt")
    with open(inp, "w") as f:
        f.write("%d\n" % len(qa))
        for a, b in zip(qa, qb):
            f.write(" ".join(repr(float(x)) for x in a) + " " + " ".join(repr(float(x)) for x in b) + "\n")
    env = dict(os.environ, CKC_MESH_PATH=PACK)
    r = subprocess.run([os.path.join(HOST, "test_dropin_rx"), repr(eps), inp, out], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = np.array([[int(x) for x in ln.split()] for ln in open(out).read().strip().split("\n")])
    def valid(q):
        ok, _ = oracle.rx_check(q, eps)
        return ok & oracle.check_angle_limits(q) & (1 - oracle.collision_state(q).astype(np.uint8))
    ok_a, _ = oracle.rx_check(qa, eps)
    va, vb = valid(qa), valid(qb)
    assert (res[:, 0] == ok_a).all() and (res[:, 1] == va).all() and (res[:, 2] == vb).all()
    want = np.array([int(va[i] and vb[i] and _rx_rbs_oracle(oracle, qa[i], qb[i], eps)) for i in range(len(qa))])
    assert (res[:, 3] == want).all() and 5 <= want.sum() < len(want)


# ---- members left untested by round 1: checkMotion, reconstructRBS, GD isValid(state), 5- / 3-argument IKproject, samplers, FK / IK ----
def _members(binary, tmp_path, edges, confs):
    inp, out = str(tmp_path / "in.txt"), str(tmp_path / "out.txt")
    os.makedirs(tmp_path / "paths", exist_ok=True)
    with open(inp, "w") as f:
        f.write("%d %d\n" % (len(edges), len(confs)))
        for a, b, k in edges:
            f.write(" ".join(repr(float(x)) for x in a) + " " + " ".join(repr(float(x)) for x in b) + " %d\n" % k)
        for q, ac, n0, n1 in confs:
            f.write(" ".join(repr(float(x)) for x in q) + " %d %d %d\n" % (ac, n0, n1))
    _run(binary, inp, out)
    res = {}
    for ln in open(out).read().strip().split("\n"):
        t = ln.split()
        res.setdefault(t[0], []).append([float(x) for x in t[1:]])
    return res


def test_pcs_members_vs_reference_goldens_and_oracle(tmp_path, oracle, gold):
    g = gold("reconstruct_rbs.npz"); h = gold("ikproject5.npz")
    nE, nC = len(g["qa"]), 300
    sel = np.concatenate([np.arange(150), 1000 + np.arange(150)])          # both active chains of the fixture
    edges = [(g["qa"][i], g["qb"][i], int(g["ik"][i])) for i in range(nE)]
    confs = [(h["q"][i], int(h["ac"][i]), int(h["ik_nn"][i][0]), int(h["ik_nn"][i][1])) for i in sel]
    res = _members("test_members_pcs", tmp_path, edges, confs)
    # reconstructRBS: bool and, for valid edges, the very matrix the reference binary returns (ordered midpoints between a and b)
    off = 0
    for i, row in enumerate(res["recon"]):
        want = g["rows"][off:off + g["nrows"][i]]; off += g["nrows"][i]
        assert int(row[1]) == int(g["ok"][i])
        if g["ok"][i]:
            assert int(row[2]) == len(want)
            assert np.abs(np.array(row[3:]).reshape(-1, 12) - want).max() < 1e-6
    assert sum(int(r[1]) for r in res["recon"]) == int(g["ok"].sum()) >= 40
    # checkMotion (fixed grid, PCS.cpp:389-432) against the restated routine, which calls the pinned isValidRBS on every grid point
    cm_o, _ = oracle.pcs_check_motion(g["qa"], g["qb"], 0, g["ik"])
    assert (np.array([r[1] for r in res["cm"]]) == cm_o).all() and 0 < cm_o.sum() < nE
    assert np.abs(np.array([r[2] for r in res["cm"]]) - np.linalg.norm(g["qa"] - g["qb"], axis=1)).max() < 1e-12     # stateDistance
    # 5-argument IKproject (neighbour's branch, active-chain swap, collision): the reference binary's own outputs
    r5 = np.array(res["ikp5"])
    assert (r5[:, 1] == h["valid"][sel]).all() and (r5[:, 2] == h["ac_out"][sel]).all() and (r5[:, 3:5] == h["ik"][sel]).all()
    v = h["valid"][sel] == 1
    assert np.abs(r5[v][:, 5:] - h["q_out"][sel][v]).max() < 1e-6 and v.sum() > 80 and (h["ac_out"][sel] != h["ac"][sel]).sum() > 10
    # 3-argument IKproject (random branch order): whatever it returns is closed on an identified branch
    r3 = np.array(res["ikp3"]); got = r3[:, 1] == 1
    idk = oracle.identify_state_ik(r3[got][:, 2:])
    assert got.sum() > 100 and (idk >= 0).all()
    for i, row in enumerate(r3[got][:20]):
        ac = confs[int(row[0])][1]
        okp, qp = oracle.pcs_project(row[2:][None], ac, idk[i][ac])
        assert okp[0] and np.abs(qp[0] - row[2:]).max() < 1e-6
    # FKsolve_rob / calc_all_IK_solutions / IsRobotsFeasible_R1
    for row, allr, fe in zip(res["fk"], res["ikall"], res["feas"]):
        i = int(row[0]); robot = 1 + (i & 1)
        T = np.array(row[1:]).reshape(4, 4)
        assert np.abs(T - oracle.fk(confs[i][0][:6], robot)).max() < 1e-9
        sols = [(s, oracle.ik(T, robot, s)) for s in range(8)]
        feas = [(s, q) for s, (ok, q) in sols if ok]
        assert int(allr[1]) == len(feas)
        a = np.array(allr[2:]).reshape(-1, 7)
        for (s, q), r in zip(feas, a):
            assert int(r[0]) == s and np.abs(r[1:] - q).max() < 1e-6
        n2 = sum(oracle.pcs_project(np.concatenate([confs[i][0][:6], np.zeros(6)])[None], 0, s)[0][0] for s in range(8))
        assert int(fe[1]) == int(n2 >= 1) and int(fe[2]) == n2
    # 50 draws of sample_q: each closed, identified on one chain at least, and accepted by the reference's rule (PCS.cpp:217)
    S = np.array(res["sample"])[:, 1:]
    assert len(S) == 50 and len(np.unique(S[:, 0])) == 50
    assert (oracle.identify_state_ik(S) >= 0).any(axis=1).all()
    assert not ((oracle.collision_state(S) == 1) & (oracle.check_angle_limits(S) == 0)).any()
    assert np.abs(np.array(res["qinv"][0]).reshape(4, 4) - np.eye(4)).max() < 1e-12
    log = open(tmp_path / "paths" / "perf_log.txt").read().split()
    assert len(log) == 16 and log[0] == "1" and float(log[2]) == 1.5 and int(log[8]) == 7          # the order of PCS.cpp:628-651


def test_gd_members_vs_oracle(tmp_path, oracle):
    from abb_dualarm_planning_b200 import sampling
    from oracle import reflibs
    q = sampling.sample_gd(61, 0, 30000)
    v, qp = oracle.gd_validate(q)
    a = qp[v == 1][:24]
    rng = np.random.default_rng(9)
    vb, b = oracle.gd_validate(a + rng.normal(size=a.shape) * np.where(np.arange(len(a))[:, None] % 3 == 0, 0.4, 0.12))
    edges = [(a[i], b[i], 0) for i in range(len(a)) if vb[i]]
    confs = [(q[i], 0, 0, 0) for i in range(120)]
    res = _members("test_members_gd", tmp_path, edges, confs)
    nvalid = 0
    for i, (row, cm) in enumerate(zip(res["recon"], res["cm"])):
        ok_o, rows_o = oracle.gd_reconstruct_rbs(edges[i][0], edges[i][1])
        assert int(row[1]) == int(ok_o)
        rbs_o, _ = oracle.gd_check_motion_rbs(edges[i][0][None], edges[i][1][None])
        cm_o, _ = oracle.gd_check_motion(edges[i][0][None], edges[i][1][None])
        assert int(cm[1]) == int(cm_o[0]) and int(cm[2]) == int(rbs_o[0])
        if ok_o:
            nvalid += 1
            got = np.array(row[3:]).reshape(-1, 12)
            assert len(got) == len(rows_o)
            d = np.abs(got - rows_o); d = np.minimum(d, np.abs(d - 2 * 3.1416))
            assert d.max() < 1e-4                                   # GD joint tolerance, see test_gd_projection_vs_oracle
            if reflibs.available():
                assert reflibs.tsb_check(got[1:-1], 1)[1].all()     # every returned midpoint closes the chain by the reference's own check
    assert nvalid >= 8
    # isValid(const ob::State*) of the GD class: projects the state in place (GD.cpp:115-134)
    iv = np.array(res["isvalid"]); valid_o, qo_o = oracle.gd_validate(q[:120])
    assert (iv[:, 1] == valid_o).all()
    d = np.abs(iv[valid_o == 1][:, 2:] - qo_o[valid_o == 1]); d = np.minimum(d, np.abs(d - 2 * 3.1416))
    assert d.max() < 1e-4 and (iv[valid_o == 0][:, 2:] == q[:120][valid_o == 0]).all()
    for row in res["fk"]:
        assert np.abs(np.array(row[1:]).reshape(4, 4) - oracle.kdl_fk(q[int(row[0])])).max() < 1e-9
    S = np.array(res["sample"])[:, 1:]
    assert len(S) == 20 and oracle.collision_state(S).sum() == 0 and oracle.check_angle_limits(S).all()
    if reflibs.available():
        assert reflibs.tsb_check(S, 1)[1].all()
    sew = res["sew"][0]
    M = np.array(sew[2:]).reshape(-1, 12)
    # literal GD.cpp:352-376: the loop variable is overwritten by the length of the step just taken, so one sewing step is made
    assert int(sew[0]) == 1 and len(M) == int(sew[1]) == 3 and np.abs(M[0] - edges[0][0]).max() == 0 and np.abs(M[-1] - edges[0][1]).max() == 0
    assert np.linalg.norm(M[1] - M[0]) <= 0.05 and oracle.collision_state(M[1][None])[0] == 0
    assert res["n"][0] == [12.0, 0.05]
    log = open(tmp_path / "paths" / "perf_log.txt").read().split()
    assert len(log) == 16 and log[0] == "1" and int(log[8]) == 7                                     # GD.cpp:447-470


def test_rx_members_vs_oracle(tmp_path, oracle, gold):
    g = gold("rx_confs_eps0.5_env2.npz")            # env-2 fixture is not usable with the env-1 driver: build env-1 inputs from the eps 0.7 set
    g = gold("rx_confs_eps0.7.npz")
    qa = g["q"][:30]
    rng = np.random.default_rng(4)
    qb = qa + rng.normal(size=qa.shape) * 0.02
    res = _members("test_members_rx", tmp_path, [(qa[i], qb[i], 0) for i in range(30)], [(qa[i], 0, 0, 0) for i in range(30)])
    eps = 0.3                                        # the class default (RX.h: epsilon = 0.3)

    def valid(x):
        ok, _ = oracle.rx_check(x, eps)
        return ok & oracle.check_angle_limits(x) & (1 - oracle.collision_state(x).astype(np.uint8))
    for i, (row, cm) in enumerate(zip(res["recon"], res["cm"])):
        want = int(valid(qa[i][None])[0] and valid(qb[i][None])[0]) and int(_rx_rbs_oracle(oracle, qa[i], qb[i], eps))
        tree_ok = int(_rx_rbs_oracle(oracle, qa[i], qb[i], eps))
        assert int(row[1]) == tree_ok and int(cm[2]) == tree_ok
        if tree_ok:
            M = np.array(row[3:]).reshape(-1, 12)
            assert (M[0] == qa[i]).all() and (M[-1] == qb[i]).all()
            assert (np.linalg.norm(np.diff(M, axis=0), axis=1) < 0.05).all()            # consecutive rows closer than RBS_tol
            t = np.linalg.norm(M - qa[i], axis=1)
            assert (np.diff(t) > 0).all()                                                # ordered along the (straight, unprojected) edge
    iv = np.array(res["isvalid"])
    assert (iv[:, 1] == valid(qa)).all()
    S = np.array(res["sample"])[:, 1:] if "sample" in res else np.zeros((0, 12))
    if len(S):
        ok, _ = oracle.rx_check(S, 0.5)
        assert ok.all() and oracle.check_angle_limits(S).all() and oracle.collision_state(S).sum() == 0
    c = res["counts"][0]
    assert c[0] == len(S) and c[0] + c[1] <= 20000 and c[2] == 0.5
