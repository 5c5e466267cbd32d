"""tests/golden/sgd_seed1_1e6_mask.npz: validity mask of the 1e6-sample S-GD seed set (seed 1) computed by the CPU restatement
(oracle/, all host cores).  NOT produced by real KDL (absent, DESIGN.md "Oracle"): it pins the device against the restatement
on the north star's seed-set size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import multiprocessing as mp
from gd_mask_agreement import work

if __name__ == "__main__":
    N, chunk = 1_000_000, 20000
    jobs = [(1, i0, min(chunk, N - i0)) for i0 in range(0, N, chunk)]
    t0 = time.time()
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        res = pool.map(work, jobs)
    vo = np.zeros(N, np.uint8)
    for i0, v, _ in res:
        vo[i0:i0 + len(v)] = v
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sgd_seed1_1e6_mask.npz")
    np.savez_compressed(out, seed=1, n=N, valid_bits=np.packbits(vo))
    print("valid %d of %d, %.1f s" % (vo.sum(), N, time.time() - t0))
