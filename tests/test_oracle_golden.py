"""The CPU oracle restatement against golden vectors produced by the GENUINE reference code
(tools/gen_golden.py: the reference's prebuilt tests/trbspcs + its fixture files).  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD

TOL_Q = 1e-6      # rad, north-star tolerance for IK joint solutions


def test_sample_generator_spec(oracle):
    # specification: u = (splitmix64(splitmix64(seed) ^ (16 i + k)) >> 11) * 2^-53 ; known answers pin the spec
    M = (1 << 64) - 1

    def sm(x):
        x = (x + 0x9E3779B97F4A7C15) & M
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)
    for seed, i, k in ((1, 0, 0), (1, 12345, 7), (8, 999999, 12), (2 ** 40 + 3, 2 ** 33, 5)):
        want = (sm(sm(seed) ^ ((16 * i + k) & M)) >> 11) * 2.0 ** -53
        assert oracle.L.ckco_uniform(seed, i, k) == want
    q, ik = oracle.sample_pcs(1, 0, 1000)
    lo = np.array([-165, -110, -110, -160, -120]) * 3.1416 / 180
    hi = np.array([165, 110, 70, 160, 120]) * 3.1416 / 180
    for r in (0, 6):
        assert (q[:, r:r + 5] >= lo).all() and (q[:, r:r + 5] <= hi).all()
        assert (np.abs(q[:, r + 5]) <= 3.1416).all()
    assert ik.min() == 0 and ik.max() == 7


def test_projection_matches_reference_binary(oracle, gold):
    g = gold("pcs_project.npz")
    for ac, okn, qn in ((0, "ok_ac0", "q_ac0"), (1, "ok_ac1", "q_ac1")):
        ok, q = oracle.pcs_project(g["q"], ac, g["ik"])
        assert (ok == g[okn]).all()
        m = ok == 1
        assert m.sum() > 1000
        assert np.abs(q[m] - g[qn][m]).max() < 1e-12          # far inside the 1e-6 rad north-star tolerance
        assert (q[~m] == g["q"][~m]).all()                     # failed projection leaves the state untouched


def test_joint_limits_and_nan_policy(oracle):
    q = np.zeros((4, 12))
    q[1, 2] = 70.01 * 3.1416 / 180
    q[2, 11] = -400.01 * 3.1416 / 180
    q[3, 5] = 399.9 * 3.1416 / 180
    assert oracle.check_angle_limits(q).tolist() == [1, 0, 0, 1]
    qn = np.zeros((1, 12)); qn[0, 3] = np.nan                 # quirk 7: NaN passes every fabs(q) > limit test
    assert oracle.check_angle_limits(qn).tolist() == [1]


def test_tridist_bit_exact_vs_pqp(oracle, gold):
    g = gold("tridist.npz")
    d = oracle.tri_dist(g["S"], g["T"])
    assert (d == g["d"]).all()                                # bit exact, including degenerate triangles and overlaps (0)


def test_collision_per_pair_flags_vs_pqp(oracle, gold):
    g = gold("collide.npz")
    flags = np.unpackbits(g["pair_flags"], axis=1)[:, :116]
    cs, pp = oracle.collision_state(g["q"], per_pair=True)
    assert (pp == flags).all()
    assert (cs == g["collision_state"]).all()
    assert (oracle.collision_state(g["q"]) == g["collision_state"]).all()     # early-exit variant


def test_hierarchy_is_conservative_vs_brute_force(oracle, gold):
    g = gold("collide.npz")
    q = g["q"][:6]
    a, pa = oracle.collision_state(q, per_pair=True)
    b, pb = oracle.collision_state(q, per_pair=True, brute=True)
    assert (pa == pb).all() and (a == b).all()


def test_spcs_million_sample_mask(oracle, gold):
    """north star: the mask on the 1e6-sample seed set must equal the reference's. The oracle is checked on a slice
    here (CPU time); tests/test_gpu_parity.py checks the CUDA path on all 1e6."""
    g = gold("spcs_seed1_1e6_mask.npz")
    n = int(g["n"])
    ok_ref = np.unpackbits(g["ok_bits"])[:n]; valid_ref = np.unpackbits(g["valid_bits"])[:n]
    assert ok_ref.sum() == 23883 and valid_ref.sum() == 11514
    i0, m = 400000, 60000
    q, ik = oracle.sample_pcs(1, i0, m)
    ok, valid, _ = oracle.pcs_validate(q, 0, ik)
    assert (ok == ok_ref[i0:i0 + m]).all()
    assert (valid == valid_ref[i0:i0 + m]).all()


def test_reference_path_fixture(oracle, gold):
    g = gold("ref_path.npz")                                   # /root/reference/paths/path.txt, 321 solved-path rows
    assert (g["collision_state"] == 0).all()
    assert (oracle.collision_state(g["path"]) == 0).all()
    # closure consistency at the 6-digit precision of the text file
    ik = oracle.identify_state_ik(g["path"])
    assert (ik >= 0).all()
    best = np.full(len(g["path"]), np.inf)
    for s in range(8):                                         # closest IK branch reproduces the passive arm
        ok, q = oracle.pcs_project(g["path"], 0, s)
        best = np.minimum(best, np.where(ok == 1, np.abs(q - g["path"]).max(axis=1), np.inf))
    assert best.max() < 2e-3
    # KDL-chain FK closes the loop on every row (kdl_class.cpp chain conventions)
    for row in g["path"][::16]:
        T = oracle.kdl_fk(row)
        assert abs(T[0, 3] - 900) < 0.5 and abs(T[1, 3]) < 0.5 and abs(T[2, 3]) < 0.5
        assert np.abs(T[:3, :3] - np.eye(3)).max() < 2e-3


def test_gd_golden_mask_prefix(oracle, gold):
    """the committed S-GD mask is what the restatement computes (first 3000 samples re-done here)"""
    from abb_dualarm_planning_b200 import sampling
    g = gold("sgd_seed1_1e6_mask.npz")
    want = np.unpackbits(g["valid_bits"])[:3000]
    v, _ = oracle.gd_validate(sampling.sample_gd(int(g["seed"]), 0, 3000))
    assert (v == want).all() and want.sum() > 20


def test_is_valid_full_vs_binary(oracle, gold):
    """isValid(const ob::State*) PCS.cpp:323-357 run through the genuine binary; unidentified branches (-1) are UB there"""
    g = gold("isvalid_full.npz")
    ident = (g["ik"] >= 0).all(1)
    v = oracle.pcs_is_valid_full(g["q"])
    assert (v[ident] == g["valid"][ident]).all() and ident.sum() > 500 and g["valid"].sum() > 300
    assert (v[~ident] == 0).all()


def test_identify_state_ik_vs_binary(oracle, gold):
    g = gold("identify_ik.npz")
    assert (oracle.identify_state_ik(g["q"]) == g["ik"]).all()


def test_rbs_vs_binary(oracle, gold):
    g = gold("rbs.npz")
    n = 120
    ok, nc = oracle.pcs_check_motion_rbs(g["qa"][:n], g["qb"][:n], 0, g["ik"][:n])
    assert (ok == g["ok"][:n]).all()
    assert (nc == g["nchecks"][:n]).all()


def test_rx_fixtures(oracle, oracle_env2, gold):
    """configs the reference's RX checker accepted (tests/results/data/rlx_rand_confs_*.txt): residual < eps"""
    for name in ("rx_confs_eps0.7.npz", "rx_confs_eps1.0.npz", "rx_confs_eps0.5_env2.npz"):
        g = gold(name)
        orc = oracle if int(g["env"]) == 1 else oracle_env2
        ok, res = orc.rx_check(g["q"], float(g["eps"]))                 # no widening: every stored configuration stays below eps when recomputed
        assert ok.all()
        assert res.max() > 0.9 * float(g["eps"])              # the accepted set fills the bound


def test_gd_projection_statistics(oracle):
    """KDL is not available: the restated Newton projection is pinned by closure + the reference's statistics
    (tests/results/kdl_gd.txt: 100 % convergence from uniform seeds, mean projection distance 7.32)."""
    q = oracle.sample_gd(1, 0, 400)
    ok, conv, it, qo = oracle.gd_project(q)
    assert conv.all()
    assert 8.0 < it.mean() < 11.5 and it.max() < 60
    for a in qo[:50]:
        T = oracle.kdl_fk(a)
        # wrapping by 2*3.1416 instead of 2*pi (kdl_class.cpp:117-121) perturbs the converged point by ~1e-5 rad
        assert np.abs(T[:3, 3] - [900, 0, 0]).max() < 0.1 and np.abs(T[:3, :3] - np.eye(3)).max() < 2e-4
    d = np.linalg.norm(qo - q, axis=1)
    assert 6.0 < d.mean() < 8.5
    assert 0.01 < ok.mean() < 0.08                            # ~3.5 % end inside the joint limits


def test_work_model_fixture():
    wm = json.load(open(os.path.join(GOLD, "work_model.json")))
    assert 0.02 < wm["f_ik"] < 0.03 and wm["n_bv_mean"] > 100


def test_env2_mask_vs_binary(oracle_env2, gold):
    g = gold("spcs_env2_seed4_60k_mask.npz")
    n = int(g["n"])
    q, ik = oracle_env2.sample_pcs(4, 0, n)
    ok, valid, _ = oracle_env2.pcs_validate(q, 0, ik)
    assert (ok == np.unpackbits(g["ok_bits"])[:n]).all() and (valid == np.unpackbits(g["valid_bits"])[:n]).all()
    assert ok.sum() == 749 and valid.sum() == 314


# ---- pins compiled from the reference's own sources (oracle/_ref/, built by oracle/Makefile where /root/reference is mounted) ----
def _reflibs():
    from oracle import reflibs
    if not reflibs.available():
        pytest.skip("oracle/_ref/ libraries not built (they are compiled from /root/reference, which is not mounted here)")
    return reflibs


def test_restatement_vs_apc_class_compiled_from_source(oracle, oracle_env2):
    """FK, closure projection and joint limits of oracle/ckc_oracle_kin.cpp against the reference's own proj_classes/apc_class.cpp
    (oracle/_ref/libapc_ref.so), both environments, both active chains"""
    rl = _reflibs()
    for env, orc in ((1, oracle), (2, oracle_env2)):
        q, ik = orc.sample_pcs(11 + env, 0, 3000)
        for robot in (0, 1):
            T = rl.apc_fk(q[:200, 6 * robot:6 * robot + 6], robot + 1, env)          # the reference numbers its robots 1 and 2
            To = np.array([orc.fk(r[6 * robot:6 * robot + 6], robot + 1) for r in q[:200]])
            assert np.abs(T - To).max() < 1e-9                                        # mm; the reference evaluates 6 kB symbolic expressions
        for ac in (0, 1):
            ok_r, q_r = rl.apc_project(q, ac, ik, env)
            ok_o, q_o = orc.pcs_project(q, ac, ik)
            assert (ok_r == ok_o).all() and ok_o.sum() > 20
            assert np.abs(q_r[ok_o == 1] - q_o[ok_o == 1]).max() < 1e-10
        qq = np.concatenate([q_o, q * 1.2])
        assert (rl.apc_limits(qq, env) == orc.check_angle_limits(qq)).all()


def test_gd_restatement_pinned_by_reference_closure_check(oracle, oracle_env2, gold):
    """The reference's own closure check (verification_class::Tsb_matrix / test_constraint compiled from source -- the test its
    tests/test_gd_kdl.cpp applies to kdl::GD) accepts every point the restated Newton projection returns, rejects the raw
    seeds, accepts the reference's solved path; the restatement's statistics sit on the reference's log (tests/results/kdl_gd.txt)"""
    rl = _reflibs()
    pin = json.load(open(os.path.join(GOLD, "gd_pin.json")))
    path = gold("ref_path.npz")["path"]
    _, ok = rl.tsb_check(path, 1)
    assert ok.all()                                               # paths/path.txt: 321 closed configurations
    assert rl.tsb_test_constraint(path[17], 1)                    # the member itself, not only the silent restatement of its rule
    from abb_dualarm_planning_b200 import sampling
    for env, orc, key in ((1, oracle, "env1"), (2, oracle_env2, "env2")):
        q = sampling.sample_gd(1, 0, 4000)
        okp, conv, it, qo = orc.gd_project(q)
        assert conv.all()                                         # reference log: 100 % success from uniform seeds
        _, passed = rl.tsb_check(qo, env)
        assert passed.all()
        _, raw = rl.tsb_check(q[:500], env)
        assert raw.sum() == 0
        hist = np.array(pin[key]["iteration_histogram"], float) / pin[key]["n"]
        mine = np.bincount(it, minlength=64)[:64] / len(it)
        assert np.abs(hist - mine).sum() < 0.08                   # 4000 samples of the fixture's 1e6-sample distribution
        d = np.linalg.norm(qo - q, axis=1).mean()
        assert abs(d - pin[key]["projection_distance_mean"]) < 0.15
    ref = pin["reference_log"]
    assert ref["kdl_success_rate"] == 1.0 and pin["env1"]["converged"] == pin["env1"]["n"]
    assert abs(pin["env1"]["projection_distance_mean"] - ref["kdl_distance_mean"]) < 0.05      # 7.34 vs the reference's 7.32 (N = 10 000)
    assert pin["env1"]["closure_check_passed"] == pin["env1"]["n"] and pin["env2"]["closure_check_passed"] == pin["env2"]["n"]
