"""measurement: agreement of the S-GD validity mask (seed 1, N samples) between the device and the CPU restatement; run on the GPU box.
usage: gd_mask_agreement.py [N=1000000] [env=1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multiprocessing as mp

def work(args):
    env, i0, m = args
    from abb_dualarm_planning_b200 import sampling
    from oracle import lib
    O = lib.Oracle(env)
    q = sampling.sample_gd(1, i0, m)
    v, qo = O.gd_validate(q)
    return i0, v, qo

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    env = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    from abb_dualarm_planning_b200 import Checker, sampling
    ck = Checker(env)
    q = sampling.sample_gd(1, 0, N)
    t0 = time.time(); valid, qo = ck.gd_validate(q); tg = time.time() - t0
    ncore = os.cpu_count() or 1
    chunk = 20000
    jobs = [(env, i0, min(chunk, N - i0)) for i0 in range(0, N, chunk)]
    t0 = time.time()
    with mp.get_context("spawn").Pool(ncore) as pool: res = pool.map(work, jobs)
    tc = time.time() - t0
    vo = np.zeros(N, np.uint8); qoo = np.zeros((N, 12))
    for i0, v, qq in res: vo[i0:i0 + len(v)] = v; qoo[i0:i0 + len(v)] = qq
    d = np.abs(qo - qoo); d = np.minimum(d, np.abs(d - 2 * 3.1416)).max(axis=1)
    print("env %d N %d: device valid %d, restatement valid %d, mask mismatches %d (device-only %d, cpu-only %d); joints > 1e-6: %d, > 1e-4: %d, > 1e-2: %d; device %.2f s, cpu %.1f s on %d cores"
          % (env, N, valid.sum(), vo.sum(), (valid != vo).sum(), ((valid == 1) & (vo == 0)).sum(), ((valid == 0) & (vo == 1)).sum(),
             (d > 1e-6).sum(), (d > 1e-4).sum(), (d > 1e-2).sum(), tg, tc, ncore))
    ck.close()
