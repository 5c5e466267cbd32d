"""ctypes binding of the CPU oracle restatement (test infrastructure, NOT product code)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libckc_oracle.so")
REPO = os.path.dirname(_HERE)
MESH_PACK = os.path.join(REPO, "abb_dualarm_planning_b200", "data", "meshes.ckcmesh")

NUM_FRAMES = 21
MAX_PAIRS = 116


def build(force=False):
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _HERE, "_build/libckc_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class Kin(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("b", "l1", "l2", "l3a", "l3b", "l4", "l5", "lee", "L", "D")] + [
        ("lim", C.c_double * 12), ("pose1", C.c_double * 3), ("pose2", C.c_double * 3)]


class Stats(C.Structure):
    _fields_ = [("bv_tests", C.c_int64), ("tri_tests", C.c_int64), ("pairs_hit", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.ckco_uniform.restype = C.c_double
        L.ckco_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.ckco_world_create_from_pack.restype = C.c_void_p
        L.ckco_world_create_from_pack.argtypes = [C.c_int, C.c_char_p]
        L.ckco_world_create_from_tris.restype = C.c_void_p
        L.ckco_world_create_from_tris.argtypes = [C.c_int, C.c_char_p]
        L.ckco_world_destroy.argtypes = [C.c_void_p]
        L.ckco_world_num_pairs.argtypes = [C.c_void_p]
        L.ckco_world_mesh_ntris.argtypes = [C.c_void_p, C.c_int]
        L.ckco_world_mesh_tris.restype = dp
        L.ckco_world_mesh_tris.argtypes = [C.c_void_p, C.c_int]
        L.ckco_world_pairs.restype = C.POINTER(C.c_int)
        L.ckco_world_pairs.argtypes = [C.c_void_p]
        L.ckco_min_pair_distance.restype = C.c_double
        L.ckco_tri_dist.restype = C.c_double
        _lib = L
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """Scalar FP64 restatement of the reference hot path; every method loops in C."""

    def __init__(self, env=1, tris_dir=None):
        self.L = lib()
        self.env = env
        self.kin = Kin()
        self.L.ckco_kin_init(C.byref(self.kin), env)
        if tris_dir is not None:
            self.world = self.L.ckco_world_create_from_tris(env, tris_dir.encode())
        else:
            self.world = self.L.ckco_world_create_from_pack(env, MESH_PACK.encode())
        if not self.world:
            raise RuntimeError("oracle: cannot load meshes")
        self.num_pairs = self.L.ckco_world_num_pairs(C.c_void_p(self.world))

    def __del__(self):
        try:
            if getattr(self, "world", None):
                self.L.ckco_world_destroy(C.c_void_p(self.world))
                self.world = None
        except Exception:
            pass

    @property
    def _w(self):
        return C.c_void_p(self.world)

    # ---- sample generators ----
    def sample_pcs(self, seed, i0, n):
        q = np.empty((n, 12)); ik = np.empty(n, np.int8); t = C.c_int()
        buf = (C.c_double * 12)()
        for i in range(n):
            self.L.ckco_sample_pcs(C.c_uint64(seed), C.c_uint64(i0 + i), buf, C.byref(t))
            q[i] = buf[:]; ik[i] = t.value
        return q, ik

    def sample_gd(self, seed, i0, n):
        q = np.empty((n, 12)); buf = (C.c_double * 12)()
        for i in range(n):
            self.L.ckco_sample_gd(C.c_uint64(seed), C.c_uint64(i0 + i), buf)
            q[i] = buf[:]
        return q

    # ---- kinematics ----
    def fk(self, q6, robot):
        T = (C.c_double * 16)(); q = np.ascontiguousarray(q6, np.float64)
        self.L.ckco_fk(C.byref(self.kin), _p(q), int(robot), T)
        return np.array(T[:]).reshape(4, 4)

    def ik(self, T, robot, sol):
        T = np.ascontiguousarray(T, np.float64).reshape(16); q = (C.c_double * 6)()
        ok = self.L.ckco_ik(C.byref(self.kin), _p(T), int(robot), int(sol), q)
        return bool(ok), np.array(q[:])

    def pcs_project(self, q, ac, ik):
        q = np.array(q, np.float64).reshape(-1, 12).copy(); n = len(q)
        ac = np.broadcast_to(np.asarray(ac, np.int8), (n,)); ik = np.broadcast_to(np.asarray(ik, np.int8), (n,))
        ok = np.zeros(n, np.uint8)
        for i in range(n):
            ok[i] = self.L.ckco_pcs_project(C.byref(self.kin), _p(q[i]), int(ac[i]), int(ik[i]))
        return ok, q

    def check_angle_limits(self, q):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 12)
        return np.array([self.L.ckco_check_angle_limits(C.byref(self.kin), _p(r)) for r in q], np.uint8)

    def identify_state_ik(self, q):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 12); out = np.empty((len(q), 2), np.int32); t = (C.c_int * 2)()
        for i, r in enumerate(q):
            self.L.ckco_identify_state_ik(C.byref(self.kin), _p(r), t); out[i] = t[:]
        return out

    # ---- collision ----
    def link_frames(self, q):
        q = np.ascontiguousarray(q, np.float64).reshape(12); fr = np.empty((NUM_FRAMES, 12))
        self.L.ckco_link_frames(self._w, _p(q), _p(fr))
        return fr

    def pairs(self):
        p = self.L.ckco_world_pairs(self._w)
        return np.ctypeslib.as_array(p, shape=(self.num_pairs, 4)).copy()

    def mesh_tris(self, m):
        n = self.L.ckco_world_mesh_ntris(self._w, m)
        if n == 0:
            return np.zeros((0, 9))
        return np.ctypeslib.as_array(self.L.ckco_world_mesh_tris(self._w, m), shape=(n, 9)).copy()

    def collision_state(self, q, per_pair=False, stats=False, brute=False):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 12); n = len(q)
        out = np.zeros(n, np.int32); pp = np.zeros((n, MAX_PAIRS), np.uint8); st = Stats()
        for i in range(n):
            if brute:
                out[i] = self.L.ckco_collision_state_brute(self._w, _p(q[i]), _p(pp[i], C.c_uint8))
            else:
                out[i] = self.L.ckco_collision_state(self._w, _p(q[i]), 0 if (per_pair or stats) else 1,
                                                     C.byref(st) if stats else None, _p(pp[i], C.c_uint8) if per_pair else None)
        res = [out]
        if per_pair: res.append(pp[:, :self.num_pairs])
        if stats: res.append((st.bv_tests, st.tri_tests, st.pairs_hit))
        return res[0] if len(res) == 1 else tuple(res)

    def min_pair_distance(self, q):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 12)
        return np.array([self.L.ckco_min_pair_distance(self._w, _p(r)) for r in q])

    def tri_dist(self, S, T):
        S = np.ascontiguousarray(S, np.float64).reshape(-1, 9); T = np.ascontiguousarray(T, np.float64).reshape(-1, 9)
        P = (C.c_double * 3)(); Q = (C.c_double * 3)()
        return np.array([self.L.ckco_tri_dist(P, Q, _p(S[i]), _p(T[i])) for i in range(len(S))])

    # ---- checker level ----
    def pcs_validate(self, q, ac, ik, stats=False):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 12); n = len(q)
        ac = np.ascontiguousarray(np.broadcast_to(np.asarray(ac, np.int8), (n,))); ik = np.ascontiguousarray(np.broadcast_to(np.asarray(ik, np.int8), (n,)))
        ok = np.zeros(n, np.uint8); valid = np.zeros(n, np.uint8); qo = np.empty((n, 12)); st = Stats()
        self.L.ckco_batch_pcs_validate(C.byref(self.kin), self._w, C.c_int64(n), _p(q), _p(ac, C.c_int8), _p(ik, C.c_int8),
                                       _p(ok, C.c_uint8), _p(valid, C.c_uint8), _p(qo), C.byref(st) if stats else None)
        if stats:
            return ok, valid, qo, (st.bv_tests, st.tri_tests, st.pairs_hit)
        return ok, valid, qo

    def pcs_is_valid_full(self, q):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 12)
        return np.array([self.L.ckco_pcs_is_valid_full(C.byref(self.kin), self._w, _p(r)) for r in q], np.uint8)

    def pcs_check_motion_rbs(self, qa, qb, ac, ik):
        qa = np.ascontiguousarray(qa, np.float64).reshape(-1, 12); qb = np.ascontiguousarray(qb, np.float64).reshape(-1, 12); n = len(qa)
        ac = np.broadcast_to(np.asarray(ac, np.int8), (n,)); ik = np.broadcast_to(np.asarray(ik, np.int8), (n,))
        ok = np.zeros(n, np.uint8); nc = np.zeros(n, np.int64); t = C.c_int64()
        for i in range(n):
            ok[i] = self.L.ckco_pcs_check_motion_rbs(C.byref(self.kin), self._w, _p(qa[i]), _p(qb[i]), int(ac[i]), int(ik[i]), C.byref(t))
            nc[i] = t.value
        return ok, nc

    # ---- GD / RX ----
    def kdl_fk(self, q):
        q = np.ascontiguousarray(q, np.float64).reshape(12); T = (C.c_double * 16)()
        self.L.ckco_kdl_fk(C.byref(self.kin), _p(q), T)
        return np.array(T[:]).reshape(4, 4)

    def gd_project(self, q):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 12); n = len(q)
        ok = np.zeros(n, np.uint8); conv = np.zeros(n, np.uint8); it = np.zeros(n, np.int32); qo = np.zeros((n, 12))
        c = C.c_int(); k = C.c_int()
        for i in range(n):
            ok[i] = self.L.ckco_gd_project(C.byref(self.kin), _p(q[i]), _p(qo[i]), C.byref(c), C.byref(k))
            conv[i] = c.value; it[i] = k.value
        return ok, conv, it, qo

    def gd_validate(self, q):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 12); n = len(q)
        valid = np.zeros(n, np.uint8); qo = np.zeros((n, 12))
        for i in range(n):
            valid[i] = self.L.ckco_gd_is_valid(C.byref(self.kin), self._w, _p(q[i]), _p(qo[i]))
        return valid, qo

    def gd_check_motion_rbs(self, qa, qb):
        qa = np.ascontiguousarray(qa, np.float64).reshape(-1, 12); qb = np.ascontiguousarray(qb, np.float64).reshape(-1, 12); n = len(qa)
        ok = np.zeros(n, np.uint8); nc = np.zeros(n, np.int64); t = C.c_int64()
        for i in range(n):
            ok[i] = self.L.ckco_gd_check_motion_rbs(C.byref(self.kin), self._w, _p(qa[i]), _p(qb[i]), C.byref(t)); nc[i] = t.value
        return ok, nc

    def rx_check(self, q, eps):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 12); n = len(q)
        ok = np.zeros(n, np.uint8); res = np.zeros(n); r = C.c_double()
        for i in range(n):
            ok[i] = self.L.ckco_rx_check(C.byref(self.kin), _p(q[i]), C.c_double(eps), C.byref(r)); res[i] = r.value
        return ok, res

    # ---- members added in round 2 (ckc_oracle_more.cpp) ----
    def pcs_reconstruct_rbs(self, qa, qb, ac, ik, cap=8192):
        qa = np.ascontiguousarray(qa, np.float64).reshape(12); qb = np.ascontiguousarray(qb, np.float64).reshape(12)
        rows = np.empty((cap, 12)); n = C.c_int()
        ok = self.L.ckco_pcs_reconstruct_rbs(C.byref(self.kin), self._w, _p(qa), _p(qb), int(ac), int(ik), _p(rows), cap, C.byref(n))
        return bool(ok), rows[:n.value].copy()

    def gd_reconstruct_rbs(self, qa, qb, cap=8192):
        qa = np.ascontiguousarray(qa, np.float64).reshape(12); qb = np.ascontiguousarray(qb, np.float64).reshape(12)
        rows = np.empty((cap, 12)); n = C.c_int()
        ok = self.L.ckco_gd_reconstruct_rbs(C.byref(self.kin), self._w, _p(qa), _p(qb), _p(rows), cap, C.byref(n))
        return bool(ok), rows[:n.value].copy()

    def pcs_check_motion(self, qa, qb, ac, ik):
        qa = np.ascontiguousarray(qa, np.float64).reshape(-1, 12); qb = np.ascontiguousarray(qb, np.float64).reshape(-1, 12); n = len(qa)
        ac = np.broadcast_to(np.asarray(ac, np.int8), (n,)); ik = np.broadcast_to(np.asarray(ik, np.int8), (n,))
        ok = np.zeros(n, np.uint8); nc = np.zeros(n, np.int32); t = C.c_int()
        for i in range(n):
            ok[i] = self.L.ckco_pcs_check_motion(C.byref(self.kin), self._w, _p(qa[i]), _p(qb[i]), int(ac[i]), int(ik[i]), C.byref(t)); nc[i] = t.value
        return ok, nc

    def gd_check_motion(self, qa, qb):
        qa = np.ascontiguousarray(qa, np.float64).reshape(-1, 12); qb = np.ascontiguousarray(qb, np.float64).reshape(-1, 12); n = len(qa)
        ok = np.zeros(n, np.uint8); nc = np.zeros(n, np.int32); t = C.c_int()
        for i in range(n):
            ok[i] = self.L.ckco_gd_check_motion(C.byref(self.kin), self._w, _p(qa[i]), _p(qb[i]), C.byref(t)); nc[i] = t.value
        return ok, nc

    def pcs_ikproject5(self, q, ac, ik_nn):
        q = np.array(q, np.float64).reshape(-1, 12).copy(); n = len(q)
        ac = np.broadcast_to(np.asarray(ac, np.int32), (n,)).copy(); nn = np.ascontiguousarray(ik_nn, np.int32).reshape(n, 2)
        valid = np.zeros(n, np.uint8); ik = np.zeros((n, 2), np.int32)
        for i in range(n):
            a = C.c_int(int(ac[i]))
            valid[i] = self.L.ckco_pcs_ikproject5(C.byref(self.kin), self._w, _p(q[i]), _p(ik[i], C.c_int), C.byref(a), _p(nn[i], C.c_int))
            ac[i] = a.value
        return valid, ac.astype(np.int8), ik, q
