"""Host-side (numpy, vectorised) implementation of the counter-based synthetic-input generator specified in DESIGN.md:

    u(seed, i, k) = (mix(mix(seed) ^ (16 i + k)) >> 11) * 2^-53,     mix = splitmix64 finaliser

S-PCS: 12 joints uniform in the reference's rand_q_ambient box (proj_classes/apc_class.cpp:643-661), ik = floor(8 u_12)
S-GD : 12 joints uniform in [-3.1416, 3.1416] (validity_checkers/StateValidityCheckerGD.cpp:60-61)
The CUDA kernels implement the same specification on the device (ckc_*_seeded_dev); this copy feeds host buffers.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
PI_ = 3.1416


def _mix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return x ^ (x >> np.uint64(31))


def uniform(seed, i0, n, k):
    """u(seed, i, k) for i in [i0, i0+n) -> float64 [n]"""
    with np.errstate(over="ignore"):
        key = _mix(np.array([seed], dtype=np.uint64))[0]
        ctr = (np.arange(i0, i0 + n, dtype=np.uint64) * np.uint64(16) + np.uint64(k)) & _M
        return (_mix(key ^ ctr) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def sample_pcs(seed, i0, n):
    d2r = PI_ / 180.0
    lo = np.array([-165 * d2r, -110 * d2r, -110 * d2r, -160 * d2r, -120 * d2r, -PI_])
    hi = np.array([165 * d2r, 110 * d2r, 70 * d2r, 160 * d2r, 120 * d2r, PI_])
    q = np.empty((n, 12))
    for j in range(12):
        q[:, j] = lo[j % 6] + uniform(seed, i0, n, j) * (hi[j % 6] - lo[j % 6])
    ik = np.minimum((uniform(seed, i0, n, 12) * 8.0).astype(np.int8), 7).astype(np.int8)
    return q, ik


def sample_gd(seed, i0, n):
    q = np.empty((n, 12))
    for j in range(12):
        q[:, j] = -PI_ + uniform(seed, i0, n, j) * (2 * PI_)
    return q
