"""ctypes bindings of the libraries oracle/Makefile compiles from the reference's OWN sources where they lie under
/root/reference (test infrastructure, NOT product code; the .so files live in oracle/_ref/, git-ignored, and travel to the
GPU box with the snapshot):

    libapc_ref.so          proj_classes/apc_class.cpp            (FKsolve_rob, calc_specific_IK_solution_R1/R2, check_angle_limits)
    libtsb_ref_env{1,2}.so validity_checkers/verification_class.cpp:116-170 (Tsb_matrix, test_constraint) -- the reference's
                           own closure check, the pin of the GD flavour (tests/test_gd_kdl.cpp:61-75 applies it to kdl::GD)
"""
import ctypes as C
import os
import numpy as np

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_libs = {}


def _load(name):
    if name not in _libs:
        p = os.path.join(_REF, name)
        _libs[name] = C.CDLL(p) if os.path.exists(p) else None
    return _libs[name]


def available():
    return all(_load(n) is not None for n in ("libapc_ref.so", "libtsb_ref_env1.so", "libtsb_ref_env2.so"))


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


# ---- verification_class: closure of a 12-joint configuration ----
def tsb_check(q, env=1):
    """returns (Tsb [n,3,4], passed [n]) of the reference's Tsb_matrix / test_constraint rule (0.1 on R, 0.5 mm on p)"""
    L = _load("libtsb_ref_env%d.so" % env)
    q = np.ascontiguousarray(q, np.float64).reshape(-1, 12); n = len(q)
    T = np.empty((n, 3, 4)); ok = np.zeros(n, np.uint8)
    L.tsbref_check(C.c_long(n), _p(q), _p(T), _p(ok, C.c_uint8))
    return T, ok


def tsb_test_constraint(q, env=1):
    """the reference's member itself (prints the matrices when it fails)"""
    L = _load("libtsb_ref_env%d.so" % env)
    q = np.ascontiguousarray(q, np.float64).reshape(12)
    return bool(L.tsbref_test_constraint(_p(q)))


# ---- apc_class ----
def apc_fk(q6, robot, env=1):
    L = _load("libapc_ref.so")
    q6 = np.ascontiguousarray(q6, np.float64).reshape(-1, 6); n = len(q6)
    rb = np.ascontiguousarray(np.broadcast_to(np.asarray(robot, np.int8), (n,)))
    T = np.empty((n, 4, 4))
    L.apcref_fk(C.c_int(env), C.c_long(n), _p(q6), _p(rb, C.c_int8), _p(T))
    return T


def apc_project(q, ac, ik, env=1):
    L = _load("libapc_ref.so")
    q = np.ascontiguousarray(q, np.float64).reshape(-1, 12); n = len(q)
    ac = np.ascontiguousarray(np.broadcast_to(np.asarray(ac, np.int8), (n,))); ik = np.ascontiguousarray(np.broadcast_to(np.asarray(ik, np.int8), (n,)))
    ok = np.zeros(n, np.uint8); qo = np.empty((n, 12))
    L.apcref_project(C.c_int(env), C.c_long(n), _p(q), _p(ac, C.c_int8), _p(ik, C.c_int8), _p(ok, C.c_uint8), _p(qo))
    return ok, qo


def apc_limits(q, env=1):
    L = _load("libapc_ref.so")
    q = np.ascontiguousarray(q, np.float64).reshape(-1, 12); n = len(q)
    ok = np.zeros(n, np.uint8)
    L.apcref_limits(C.c_int(env), C.c_long(n), _p(q), _p(ok, C.c_uint8))
    return ok
