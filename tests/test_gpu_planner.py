"""GPU: the lock-step CBiRRT-PCS ensemble (abb_dualarm_planning_b200/host/cbirrt_pcs_batch.cpp) -- every
returned path is re-validated with the CPU oracle: end points, closure + IK branches of every milestone, no collision,
and every edge accepted by the oracle's checkMotionRBS with the arguments the planner recorded."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu
HOST = os.path.join(REPO, "abb_dualarm_planning_b200", "host")
PACK = os.path.join(REPO, "abb_dualarm_planning_b200", "data", "meshes.ckcmesh")
START = np.array([0.5236, 1.7453, -1.8326, -1.4835, 1.5708, 0, 1.004278, 0.2729, 0.9486, -1.15011, 1.81001, -1.97739])   # run/plan_PCS.cpp:269
GOAL = np.array([0.5236, 0.34907, 0.69813, -1.3963, 1.5708, 0, 0.7096, 1.8032, -1.7061, -1.6286, 1.9143, -2.0155])      # :271


def _plan(tmp_path, n, step, seed, flavour="pcs", workers=1, window=None):
    out = str(tmp_path / ("paths_%s_%d_%s.txt" % (flavour, n, window)))
    env = dict(os.environ, CKC_MESH_PATH=PACK)
    if window is not None:
        env["CKC_PLAN_WINDOW"] = str(window)
    r = subprocess.run([os.path.join(HOST, "cbirrt_pcs_batch"), "1", str(n), str(step), str(seed), out, str(workers), flavour], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    paths, summary = [], None
    lines = open(out).read().strip().split("\n")
    i = 0
    while i < len(lines):
        t = lines[i].split()
        if t[0] == "summary":
            summary = dict(zip(t[1::2], t[2::2])); i += 1
        else:
            m = int(t[3])
            paths.append(np.array([[float(x) for x in ln.split()] for ln in lines[i + 1:i + 1 + m]])); i += 1 + m
    return paths, summary, r.stdout


def test_cbirrt_pcs_ensemble_paths_are_valid(tmp_path, oracle):
    paths, summary, _ = _plan(tmp_path, 24, 1.0, 5)
    assert int(summary["solved"]) == 24 == len(paths)
    n_edges = 0
    for P in paths:
        q = P[:, :12]; ikn = P[:, 13:15].astype(int); e = P[:, 15:18].astype(int)
        assert np.abs(q[0] - START).max() == 0 and np.abs(q[-1] - GOAL).max() == 0
        assert (oracle.collision_state(q) == 0).all()
        assert (oracle.identify_state_ik(q) == ikn).all() and (ikn >= 0).all()      # closed on both chains, branches as recorded
        a = np.where(e[:-1, 2:3] == 0, q[:-1], q[1:]); b = np.where(e[:-1, 2:3] == 0, q[1:], q[:-1])
        ok, _ = oracle.pcs_check_motion_rbs(a, b, e[:-1, 0].astype(np.int8), e[:-1, 1].astype(np.int8))
        assert (ok == 1).all() and (e[-1] == -1).all()
        n_edges += len(a)
    assert n_edges >= 24


def test_cbirrt_single_query_matches_its_ensemble_run(tmp_path):
    """a query's random stream is its own: query 0 of an ensemble of 8 returns the path it returns when run alone"""
    p1, _, _ = _plan(tmp_path, 1, 1.0, 9)
    p8, _, _ = _plan(tmp_path, 8, 1.0, 9)
    assert p1[0].shape == p8[0].shape and np.abs(p1[0] - p8[0]).max() == 0


def test_cbirrt_window_batching_returns_the_step_by_step_planner_paths(tmp_path):
    """batching INSIDE one query: a growTree call's run of steps asked as one window (up to 64 steps per device
    round trip) must build exactly the tree the reference's one-question-at-a-time order builds (window = 1): same paths, bit for
    bit, same milestones per query -- in fewer round trips"""
    p1, s1, _ = _plan(tmp_path, 12, 1.0, 21, window=1)
    p16, s16, _ = _plan(tmp_path, 12, 1.0, 21, window=64)
    assert int(s1["solved"]) == 12 == int(s16["solved"]) and len(p1) == len(p16) == 12
    for a, b in zip(p1, p16):
        assert a.shape == b.shape and np.abs(a - b).max() == 0
    assert s1["mean_tree_nodes"] == s16["mean_tree_nodes"] and s1["mean_path_nodes"] == s16["mean_path_nodes"]
    assert float(s16["mean_rounds"]) < 0.6 * float(s1["mean_rounds"])       # measured: ~4x fewer device rounds per query


def test_cbirrt_gd_ensemble_paths_are_valid(tmp_path, oracle):
    """the GD flavour (planners/CBiRRT_GD.cpp): kdl::GD projection + RBS through ckc_gd_validate / ckc_gd_check_motion_rbs"""
    paths, summary, _ = _plan(tmp_path, 12, 2.0, 3, "gd", workers=2)
    assert int(summary["solved"]) == 12 == len(paths)
    for P in paths:
        q = P[:, :12]; e = P[:, 15:18].astype(int)
        assert np.abs(q[0] - START).max() == 0 and np.abs(q[-1] - GOAL).max() == 0
        assert (oracle.collision_state(q) == 0).all()
        for row in q[1:-1]:                                   # projected milestones close the KDL chain (reference tolerance 0.1 mm)
            T = oracle.kdl_fk(row)
            assert np.abs(T[:3, 3] - [900, 0, 0]).max() < 0.1 and np.abs(T[:3, :3] - np.eye(3)).max() < 2e-4
        a = np.where(e[:-1, 2:3] == 0, q[:-1], q[1:]); b = np.where(e[:-1, 2:3] == 0, q[1:], q[:-1])
        ok, _ = oracle.gd_check_motion_rbs(a, b)
        assert (ok == 1).all()


def test_cbirrt_hb_ensemble_paths_are_valid(tmp_path, oracle):
    """the hybrid flavour (planners/CBiRRT_HB.cpp): kdl::GD projection of every step, both IK branches identified, APC bisection as local
    connection -- milestones closed on both chains with the recorded branches, collision free, every edge accepted by the oracle's PCS
    checkMotionRBS with the recorded arguments"""
    paths, summary, _ = _plan(tmp_path, 12, 1.0, 11, "hb", workers=2)
    assert int(summary["solved"]) == 12 == len(paths) and summary["flavour"] == "hb"
    for P in paths:
        q = P[:, :12]; ikn = P[:, 13:15].astype(int); e = P[:, 15:18].astype(int)
        assert np.abs(q[0] - START).max() == 0 and np.abs(q[-1] - GOAL).max() == 0
        assert (oracle.collision_state(q) == 0).all()
        assert (oracle.identify_state_ik(q) == ikn).all()
        a = np.where(e[:-1, 2:3] == 0, q[:-1], q[1:]); b = np.where(e[:-1, 2:3] == 0, q[1:], q[:-1])
        ok, _ = oracle.pcs_check_motion_rbs(a, b, e[:-1, 0].astype(np.int8), e[:-1, 1].astype(np.int8))
        assert (ok == 1).all()


def test_prm_pcs_roadmap_path_is_valid(tmp_path, oracle):
    """the roadmap flavour (planners/PRM_PCS.cpp): batched sample_q + batched addMilestone connections"""
    out = str(tmp_path / "prm.txt")
    env = dict(os.environ, CKC_MESH_PATH=PACK)
    r = subprocess.run([os.path.join(HOST, "prm_pcs_batch"), "1", "400", "11", out], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-1500:]
    lines = open(out).read().strip().split("\n")
    m = int(lines[0].split()[3])
    P = np.array([[float(x) for x in ln.split()] for ln in lines[1:1 + m]])
    summary = dict(zip(lines[-1].split()[1::2], lines[-1].split()[2::2]))
    assert int(summary["solved"]) == 1 and int(summary["edges_valid"]) > 50
    q = P[:, :12]; ikn = P[:, 13:15].astype(int); e = P[:, 15:18].astype(int)
    assert np.abs(q[0] - START).max() == 0 and np.abs(q[-1] - GOAL).max() == 0
    assert (oracle.collision_state(q) == 0).all()
    assert (oracle.identify_state_ik(q) == ikn).all()
    a = np.where(e[:-1, 2:3] == 0, q[:-1], q[1:]); b = np.where(e[:-1, 2:3] == 0, q[1:], q[:-1])
    ok, _ = oracle.pcs_check_motion_rbs(a, b, e[:-1, 0].astype(np.int8), e[:-1, 1].astype(np.int8))
    assert (ok == 1).all()


def _plan_lazy(tmp_path, planner, n, step, seed, workers=1):
    out = str(tmp_path / ("paths_%s_%d.txt" % (planner, n)))
    env = dict(os.environ, CKC_MESH_PATH=PACK)
    r = subprocess.run([os.path.join(HOST, "lazy_gd_batch"), planner, "1", str(n), str(step), str(seed), out, str(workers)], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    paths, summary = [], None
    lines = open(out).read().strip().split("\n")
    i = 0
    while i < len(lines):
        t = lines[i].split()
        if t[0] == "summary":
            summary = dict(zip(t[1::2], t[2::2])); i += 1
        else:
            m = int(t[3])
            paths.append(np.array([[float(x) for x in ln.split()] for ln in lines[i + 1:i + 1 + m]])); i += 1 + m
    return paths, summary


@pytest.mark.parametrize("planner,step,n", [("rrt", 1.0, 6), ("lazyrrt", 1.0, 6), ("sbl", 0.6, 10)])
def test_gd_single_tree_and_lazy_planners(tmp_path, oracle, planner, step, n):
    """planners/RRT_GD.cpp, LazyRRT_GD.cpp, SBL_GD.cpp restated over the C ABI (host/lazy_gd_batch.cpp): every returned path starts at the
    start state, ends at the goal, its milestones close the KDL chain and are collision free, and every edge passes the oracle's
    checkMotionRBS in the direction the planner checked it"""
    paths, summary = _plan_lazy(tmp_path, planner, n, step, 3, workers=2)
    assert int(summary["solved"]) == n == len(paths)
    for P in paths:
        q = P[:, :12]; e = P[:, 17].astype(int)
        assert np.abs(q[0] - START).max() == 0 and np.abs(q[-1] - GOAL).max() == 0 and len(q) >= 2
        assert (oracle.collision_state(q) == 0).all()
        for row in q[1:-1]:
            T = oracle.kdl_fk(row)
            assert np.abs(T[:3, 3] - [900, 0, 0]).max() < 0.1 and np.abs(T[:3, :3] - np.eye(3)).max() < 2e-4
        a = np.where(e[:-1, None] == 0, q[:-1], q[1:]); b = np.where(e[:-1, None] == 0, q[1:], q[:-1])
        keep = np.linalg.norm(a - b, axis=1) > 0                 # SBL's junction row repeats the other tree's state
        ok, _ = oracle.gd_check_motion_rbs(a[keep], b[keep])
        assert (ok == 1).all() and e[-1] == -1


def test_lazy_planner_single_query_matches_its_ensemble_run(tmp_path):
    p1, _ = _plan_lazy(tmp_path, "sbl", 1, 0.6, 9)
    p4, _ = _plan_lazy(tmp_path, "sbl", 4, 0.6, 9, workers=1)
    assert p1[0].shape == p4[0].shape and np.abs(p1[0] - p4[0]).max() == 0
