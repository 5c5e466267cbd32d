"""ctypes binding of libckc.so (include/ckc.h).  No compute happens here: every method forwards to the C ABI."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MESH_PACK = os.path.join(_HERE, "data", "meshes.ckcmesh")
MAX_PAIRS = 116


class CkcError(RuntimeError):
    pass


def lib_path():
    """csrc/libckc.so; CKC_LIB names another build of the same library (kernel experiments on the GPU box)"""
    return os.environ.get("CKC_LIB") or os.path.join(_HERE, "csrc", "libckc.so")


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("ik_calls", "ik_feasible", "collision_calls", "collisions", "is_valid_calls", "bv_tests",
                                         "tri_tests", "queue_overflows", "gd_iterations", "kernel_launches")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def load_library():
    """Load csrc/libckc.so.  Raises CkcError when it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise CkcError("libckc.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C %s`"
                       % (p, os.path.dirname(p)))
    L = C.CDLL(p)
    vp, i64, dp = C.c_void_p, C.c_int64, C.c_void_p
    L.ckc_create.argtypes = [C.POINTER(vp), C.c_int, C.c_char_p, C.c_int]
    L.ckc_create_multi.argtypes = [C.POINTER(vp), C.c_int, C.c_char_p, C.POINTER(C.c_int), C.c_int]
    L.ckc_num_devices.argtypes = [vp]
    L.ckc_fk.argtypes = [vp, i64, dp, dp, dp]
    L.ckc_ik.argtypes = [vp, i64, dp, dp, dp, dp, dp]
    L.ckc_kdl_fk.argtypes = [vp, i64, dp, dp]
    L.ckc_pcs_reconstruct_rbs.argtypes = [vp, dp, dp, C.c_int, C.c_int, dp, i64, C.POINTER(C.c_int64), dp]
    L.ckc_gd_reconstruct_rbs.argtypes = [vp, dp, dp, dp, i64, C.POINTER(C.c_int64), dp]
    L.ckc_destroy.argtypes = [vp]; L.ckc_destroy.restype = None
    L.ckc_last_error.argtypes = [vp]; L.ckc_last_error.restype = C.c_char_p
    L.ckc_create_error.restype = C.c_char_p
    L.ckc_version.restype = C.c_char_p
    L.ckc_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.ckc_reset_stats.argtypes = [vp]
    L.ckc_num_pairs.argtypes = [vp]
    L.ckc_synchronize.argtypes = [vp]
    L.ckc_host_alloc.argtypes = [C.POINTER(vp), C.c_uint64]
    L.ckc_host_free.argtypes = [vp]
    L.ckc_pcs_project.argtypes = [vp, i64, dp, dp, dp, dp, dp]
    L.ckc_pcs_validate.argtypes = [vp, i64, dp, dp, dp, dp, dp, dp]
    L.ckc_identify_ik.argtypes = [vp, i64, dp, dp]
    L.ckc_pcs_is_valid.argtypes = [vp, i64, dp, dp]
    L.ckc_pcs_validate_compact.argtypes = [vp, i64, dp, dp, dp, dp, C.POINTER(C.c_int64), dp, dp, i64]
    L.ckc_gd_validate_compact.argtypes = [vp, i64, dp, dp, C.POINTER(C.c_int64), dp, dp, i64]
    L.ckc_check_angle_limits.argtypes = [vp, i64, dp, dp]
    L.ckc_collide.argtypes = [vp, i64, dp, dp]
    L.ckc_collide_pairs.argtypes = [vp, i64, dp, dp]
    L.ckc_pcs_validate_dev.argtypes = [vp, i64, dp, dp, dp, dp, dp, dp, vp]
    L.ckc_spcs_validate_seeded_dev.argtypes = [vp, C.c_uint64, C.c_uint64, i64, dp, dp, dp, vp]
    L.ckc_measure_fma_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.ckc_set_stage_timing.argtypes = [vp, C.c_int]
    L.ckc_set_overlap.argtypes = [vp, C.c_int]
    L.ckc_get_stage_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    for name, args in (("ckc_pcs_check_motion_rbs", [vp, i64, dp, dp, dp, dp, dp, dp]),
                       ("ckc_gd_project", [vp, i64, dp, dp, dp, dp, dp]),
                       ("ckc_gd_validate", [vp, i64, dp, dp, dp]),
                       ("ckc_gd_check_motion_rbs", [vp, i64, dp, dp, dp, dp]),
                       ("ckc_rx_validate", [vp, i64, dp, C.c_double, dp, dp]),
                       ("ckc_gd_validate_dev", [vp, i64, dp, dp, dp, vp]),
                       ("ckc_sgd_validate_seeded_dev", [vp, C.c_uint64, C.c_uint64, i64, dp, dp, dp, vp])):
        if hasattr(L, name):
            getattr(L, name).argtypes = args
    _lib = L
    return L


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)            # raw address (torch .data_ptr())


def _q(q):
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 12)
    return q, len(q)


def _i8(a, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.int8), (n,)))


class Checker:
    """One libckc context (one CUDA device).  Mirrors the batch forms of the reference's StateValidityChecker API."""

    def __init__(self, env=1, mesh_path=None, device=0, devices=None):
        """device: one CUDA ordinal; devices = [d0, d1, ...]: a multi-device context (ckc_create_multi) whose host entry points
        split every batch into contiguous ranges, one per device"""
        self.L = load_library()
        self.ctx = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.L.ckc_create_multi(C.byref(self.ctx), int(env), (mesh_path or MESH_PACK).encode(), arr, len(devices))
            device = devices[0]
        else:
            rc = self.L.ckc_create(C.byref(self.ctx), int(env), (mesh_path or MESH_PACK).encode(), int(device))
        if rc != 0:
            msg = self.L.ckc_create_error().decode()
            self.ctx = None
            raise CkcError("ckc_create failed (%d): %s" % (rc, msg))
        self.env = env
        self.device = device
        self.num_pairs = self.L.ckc_num_pairs(self.ctx)

    def close(self):
        if getattr(self, "ctx", None):
            self.L.ckc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise CkcError("libckc error %d: %s" % (rc, self.L.ckc_last_error(self.ctx).decode()))

    # ---- host-buffer API (numpy) ----
    def pcs_project(self, q, ac, ik):
        q, n = _q(q); ac = _i8(ac, n); ik = _i8(ik, n)
        qo = np.empty_like(q); ok = np.zeros(n, np.uint8)
        self._chk(self.L.ckc_pcs_project(self.ctx, n, _ptr(q), _ptr(ac), _ptr(ik), _ptr(qo), _ptr(ok)))
        return ok, qo

    def pcs_validate(self, q, ac, ik, want_q=True):
        q, n = _q(q); ac = _i8(ac, n); ik = _i8(ik, n)
        qo = np.empty_like(q) if want_q else None; ok = np.zeros(n, np.uint8); valid = np.zeros(n, np.uint8)
        self._chk(self.L.ckc_pcs_validate(self.ctx, n, _ptr(q), _ptr(ac), _ptr(ik), _ptr(qo), _ptr(ok), _ptr(valid)))
        return ok, valid, qo

    def identify_ik(self, q):
        q, n = _q(q); out = np.zeros((n, 2), np.int8)
        self._chk(self.L.ckc_identify_ik(self.ctx, n, _ptr(q), _ptr(out)))
        return out

    def pcs_is_valid(self, q):
        q, n = _q(q); valid = np.zeros(n, np.uint8)
        self._chk(self.L.ckc_pcs_is_valid(self.ctx, n, _ptr(q), _ptr(valid)))
        return valid

    def check_angle_limits(self, q):
        q, n = _q(q); ok = np.zeros(n, np.uint8)
        self._chk(self.L.ckc_check_angle_limits(self.ctx, n, _ptr(q), _ptr(ok)))
        return ok

    def collide(self, q):
        q, n = _q(q); hit = np.zeros(n, np.uint8)
        self._chk(self.L.ckc_collide(self.ctx, n, _ptr(q), _ptr(hit)))
        return hit

    def collide_pairs(self, q):
        q, n = _q(q); flags = np.zeros((n, MAX_PAIRS), np.uint8)
        self._chk(self.L.ckc_collide_pairs(self.ctx, n, _ptr(q), _ptr(flags)))
        return flags[:, :self.num_pairs]

    def pcs_check_motion_rbs(self, qa, qb, ac, ik):
        qa, n = _q(qa); qb, _ = _q(qb); ac = _i8(ac, n); ik = _i8(ik, n)
        ok = np.zeros(n, np.uint8); nc = np.zeros(n, np.int32)
        self._chk(self.L.ckc_pcs_check_motion_rbs(self.ctx, n, _ptr(qa), _ptr(qb), _ptr(ac), _ptr(ik), _ptr(ok), _ptr(nc)))
        return ok, nc

    def gd_project(self, q):
        q, n = _q(q); qo = np.empty_like(q); ok = np.zeros(n, np.uint8); conv = np.zeros(n, np.uint8); it = np.zeros(n, np.int32)
        self._chk(self.L.ckc_gd_project(self.ctx, n, _ptr(q), _ptr(qo), _ptr(ok), _ptr(conv), _ptr(it)))
        return ok, conv, it, qo

    def gd_validate(self, q):
        q, n = _q(q); qo = np.empty_like(q); valid = np.zeros(n, np.uint8)
        self._chk(self.L.ckc_gd_validate(self.ctx, n, _ptr(q), _ptr(qo), _ptr(valid)))
        return valid, qo

    def gd_validate_compact(self, q, cap=None):
        q, n = _q(q); cap = cap or max(n // 8, 1024)
        valid = np.zeros(n, np.uint8); idx = np.zeros(cap, np.int32); qv = np.zeros((cap, 12)); nv = C.c_int64()
        self._chk(self.L.ckc_gd_validate_compact(self.ctx, n, _ptr(q), _ptr(valid), C.byref(nv), _ptr(idx), _ptr(qv), cap))
        return valid, idx[:nv.value], qv[:nv.value]

    def pcs_validate_compact(self, q, ac, ik, cap=None):
        q, n = _q(q); ac = _i8(ac, n); ik = _i8(ik, n); cap = cap or max(n // 8, 1024)
        valid = np.zeros(n, np.uint8); idx = np.zeros(cap, np.int32); qv = np.zeros((cap, 12)); nv = C.c_int64()
        self._chk(self.L.ckc_pcs_validate_compact(self.ctx, n, _ptr(q), _ptr(ac), _ptr(ik), _ptr(valid), C.byref(nv), _ptr(idx), _ptr(qv), cap))
        return valid, idx[:nv.value], qv[:nv.value]

    def gd_check_motion_rbs(self, qa, qb):
        qa, n = _q(qa); qb, _ = _q(qb); ok = np.zeros(n, np.uint8); nc = np.zeros(n, np.int32)
        self._chk(self.L.ckc_gd_check_motion_rbs(self.ctx, n, _ptr(qa), _ptr(qb), _ptr(ok), _ptr(nc)))
        return ok, nc

    def rx_validate(self, q, eps):
        q, n = _q(q); valid = np.zeros(n, np.uint8); res = np.zeros(n)
        self._chk(self.L.ckc_rx_validate(self.ctx, n, _ptr(q), float(eps), _ptr(valid), _ptr(res)))
        return valid, res

    # ---- single-arm kinematics, one to one with the reference's members ----
    def fk(self, q6, robot):
        """FKsolve_rob(q, robot_num): robot numbers are the reference's 1 and 2; returns [n, 4, 4]"""
        q6 = np.ascontiguousarray(q6, np.float64).reshape(-1, 6); n = len(q6); rb = _i8(robot, n); T = np.empty((n, 4, 4))
        self._chk(self.L.ckc_fk(self.ctx, n, _ptr(q6), _ptr(rb), _ptr(T)))
        return T

    def ik(self, T, robot, sol):
        """IKsolve_rob(T, robot_num, solution_num) -> (ok [n], q6 [n, 6])"""
        T = np.ascontiguousarray(T, np.float64).reshape(-1, 16); n = len(T); rb = _i8(robot, n); so = _i8(sol, n)
        q6 = np.empty((n, 6)); ok = np.zeros(n, np.uint8)
        self._chk(self.L.ckc_ik(self.ctx, n, _ptr(T), _ptr(rb), _ptr(so), _ptr(q6), _ptr(ok)))
        return ok, q6

    def kdl_fk(self, q):
        """kdl::FK(q): tip frame of the 12-joint chain, [n, 4, 4]"""
        q, n = _q(q); T = np.empty((n, 4, 4))
        self._chk(self.L.ckc_kdl_fk(self.ctx, n, _ptr(q), _ptr(T)))
        return T

    def pcs_reconstruct_rbs(self, qa, qb, ac, ik, cap=4096):
        """reconstructRBS of one edge: (ok, ordered configurations [m, 12] from qa to qb)"""
        qa = np.ascontiguousarray(qa, np.float64).reshape(12); qb = np.ascontiguousarray(qb, np.float64).reshape(12)
        out = np.empty((cap, 12)); m = C.c_int64(); ok = np.zeros(1, np.uint8)
        self._chk(self.L.ckc_pcs_reconstruct_rbs(self.ctx, _ptr(qa), _ptr(qb), int(ac), int(ik), _ptr(out), cap, C.byref(m), _ptr(ok)))
        return bool(ok[0]), out[:m.value].copy()

    def gd_reconstruct_rbs(self, qa, qb, cap=4096):
        qa = np.ascontiguousarray(qa, np.float64).reshape(12); qb = np.ascontiguousarray(qb, np.float64).reshape(12)
        out = np.empty((cap, 12)); m = C.c_int64(); ok = np.zeros(1, np.uint8)
        self._chk(self.L.ckc_gd_reconstruct_rbs(self.ctx, _ptr(qa), _ptr(qb), _ptr(out), cap, C.byref(m), _ptr(ok)))
        return bool(ok[0]), out[:m.value].copy()

    @property
    def num_devices(self):
        return self.L.ckc_num_devices(self.ctx)

    # ---- raw-pointer API (host or device addresses, e.g. torch .data_ptr()) ----
    def pcs_validate_ptr(self, n, q, ac, ik, q_out, ok, valid):
        self._chk(self.L.ckc_pcs_validate(self.ctx, int(n), _ptr(q), _ptr(ac), _ptr(ik), _ptr(q_out), _ptr(ok), _ptr(valid)))

    def pcs_validate_dev(self, n, d_q, d_ac, d_ik, d_q_out, d_ok, d_valid, stream=0):
        self._chk(self.L.ckc_pcs_validate_dev(self.ctx, int(n), _ptr(d_q), _ptr(d_ac), _ptr(d_ik), _ptr(d_q_out), _ptr(d_ok), _ptr(d_valid), C.c_void_p(stream)))

    def spcs_validate_seeded_dev(self, seed, i0, n, d_q_out, d_ok, d_valid, stream=0):
        self._chk(self.L.ckc_spcs_validate_seeded_dev(self.ctx, int(seed), int(i0), int(n), _ptr(d_q_out), _ptr(d_ok), _ptr(d_valid), C.c_void_p(stream)))

    def gd_validate_dev(self, n, d_q, d_q_out, d_valid, stream=0):
        self._chk(self.L.ckc_gd_validate_dev(self.ctx, int(n), _ptr(d_q), _ptr(d_q_out), _ptr(d_valid), C.c_void_p(stream)))

    def sgd_validate_seeded_dev(self, seed, i0, n, d_q_out, d_ok, d_valid, stream=0):
        self._chk(self.L.ckc_sgd_validate_seeded_dev(self.ctx, int(seed), int(i0), int(n), _ptr(d_q_out), _ptr(d_ok), _ptr(d_valid), C.c_void_p(stream)))

    def synchronize(self):
        self._chk(self.L.ckc_synchronize(self.ctx))

    def stats(self):
        s = Stats()
        self._chk(self.L.ckc_get_stats(self.ctx, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._chk(self.L.ckc_reset_stats(self.ctx))

    def set_overlap(self, on):
        self._chk(self.L.ckc_set_overlap(self.ctx, 1 if on else 0))

    def set_stage_timing(self, on):
        self._chk(self.L.ckc_set_stage_timing(self.ctx, 1 if on else 0))

    def stage_times(self):
        ms = (C.c_double * 4)(); n = (C.c_int64 * 4)()
        self._chk(self.L.ckc_get_stage_times(self.ctx, ms, n))
        names = ("project", "frames", "collide", "rbs")
        return {names[i]: {"ms": ms[i], "launches": int(n[i])} for i in range(4)}

    def fma_peak(self, precision):
        t = C.c_double()
        self._chk(self.L.ckc_measure_fma_peak(self.ctx, int(precision), C.byref(t)))
        return t.value
