"""tests/golden/gd_pin.json -- what the REFERENCE holds about its GD (KDL Newton) projection, plus the CPU restatement's
statistics on the S-GD seed sets measured by the reference's own closure check.

Run in the build container (needs /root/reference for oracle/_ref/libtsb_ref_env*.so and tests/results/kdl_gd.txt):
    python tools/gen_golden_gd_pin.py

Pins recorded
  reference_log : tests/results/kdl_gd.txt (10 000 runs of the reference's tests/test_gd_kdl.cpp): KDL success rate, mean and
                  percentiles of the projection distance |q_projected - q_seed|_2
  env1 / env2   : S-GD seed 1, N samples, projected by oracle/ckc_oracle_kin.cpp (restated kdl::GD): convergence count, Newton
                  iteration histogram, mean projection distance, and the number of converged points that pass the reference's
                  verification_class::Tsb_matrix / test_constraint (compiled from the reference source) with the largest
                  rotation / position residual seen
The GPU test (tests/test_gpu_parity.py::test_gd_pinned_by_reference_closure_check) holds the device to the same numbers.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
import numpy as np


def work(args):
    env, i0, m = args
    from abb_dualarm_planning_b200 import sampling
    from oracle import lib, reflibs
    O = lib.Oracle(env)
    q = sampling.sample_gd(1, i0, m)
    ok, conv, it, qo = O.gd_project(q)
    c = conv == 1
    T, passed = reflibs.tsb_check(qo[c], env)
    D = 900.0 if env == 1 else 1200.0
    Tsd = np.array([[1, 0, 0, D], [0, 1, 0, 0], [0, 0, 1, 0]], float)
    err = np.abs(T - Tsd)
    dist = np.linalg.norm(qo[c] - q[c], axis=1)
    return (int(c.sum()), int(ok.sum()), np.bincount(it, minlength=64)[:64], float(dist.sum()), int(passed.sum()),
            float(err[:, :, :3].max()), float(err[:, :, 3].max()))


def run(env, N):
    chunk = 20000
    jobs = [(env, i0, min(chunk, N - i0)) for i0 in range(0, N, chunk)]
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        res = pool.map(work, jobs)
    hist = sum(r[2] for r in res)
    nconv = sum(r[0] for r in res)
    return {"seed": 1, "n": N, "converged": nconv, "within_joint_limits": sum(r[1] for r in res), "iteration_histogram": [int(x) for x in hist],
            "iterations_mean": float((hist * np.arange(64)).sum() / N), "projection_distance_mean": sum(r[3] for r in res) / nconv,
            "closure_check_passed": sum(r[4] for r in res), "closure_max_rotation_residual": max(r[5] for r in res),
            "closure_max_position_residual_mm": max(r[6] for r in res)}


if __name__ == "__main__":
    t0 = time.time()
    a = np.loadtxt("/root/reference/tests/results/kdl_gd.txt")
    out = {"reference_log": {"file": "tests/results/kdl_gd.txt", "columns": "kdl_success kdl_time_s kdl_distance gd_success gd_time_s gd_distance",
                             "runs": int(a.shape[0]), "kdl_success_rate": float(a[:, 0].mean()), "kdl_distance_mean": float(a[:, 2].mean()),
                             "kdl_distance_std": float(a[:, 2].std()), "kdl_distance_p01_p50_p99": [float(x) for x in np.percentile(a[:, 2], [1, 50, 99])],
                             "kdl_time_mean_s": float(a[:, 1].mean())},
           "closure_rule": "verification_class::test_constraint (verification_class.cpp:116-139): |Tsb - Tsd| <= 0.1 on the rotation part, <= 0.5 mm on the position",
           "env1": run(1, 1_000_000), "env2": run(2, 300_000)}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gd_pin.json")
    json.dump(out, open(p, "w"), indent=1)
    print(json.dumps({k: (v if k == "reference_log" else {kk: vv for kk, vv in v.items() if kk != "iteration_histogram"}) for k, v in out.items() if k != "closure_rule"}, indent=1))
    print("%.0f s" % (time.time() - t0))
