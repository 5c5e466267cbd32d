"""Driver of the GENUINE reference code inside the reference's prebuilt tests/trbspcs (oracle infrastructure only).

oracle/Makefile target `ref` places the binary, the mesh data it loads, two stub libraries and the LD_PRELOAD request
server (oracle/refshim.cpp) under oracle/_ref/.  Each call runs the binary once on a request file.
"""
import os
import struct
import subprocess
import tempfile
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(_HERE, "_ref")
OPS = dict(pcs_project=1, collide=2, pcs_validate=3, rbs=4, tridist=5, pair_counters=6, identify_ik=7, min_distance=8,
           fk=9, ik=10, spcs_seeded=11, isvalid_full=12, reconstruct_rbs=13, ikproject5=14)
MAX_PAIRS = 116


def available():
    return all(os.path.exists(os.path.join(REF, p)) for p in
               ("tests/trbspcs", "librefshim.so", "libGL.so.1", "libompl.so.12", "simulator/obs.tris"))


def _run(op, n, env, flags, payload, timeout=None):
    with tempfile.TemporaryDirectory() as td:
        req, out = os.path.join(td, "req.bin"), os.path.join(td, "out.bin")
        with open(req, "wb") as f:
            f.write(b"CKCREQ1\0" + struct.pack("<iiii", OPS[op], n, env, flags) + payload)
        e = dict(os.environ, LD_PRELOAD=os.path.join(REF, "librefshim.so"), LD_LIBRARY_PATH=REF + ":" + os.environ.get("LD_LIBRARY_PATH", ""),
                 CKC_REF_REQ=req, CKC_REF_OUT=out)
        r = subprocess.run(["./trbspcs"], cwd=os.path.join(REF, "tests"), env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
        if r.returncode != 0 or not os.path.exists(out):
            raise RuntimeError("refshim failed rc=%d: %s" % (r.returncode, r.stderr.decode()[-400:]))
        with open(out, "rb") as f:
            data = f.read()
    assert data[:7] == b"CKCRSP1"
    secs = struct.unpack("<d", data[16:24])[0]
    return data[24:], secs


def _q(q):
    return np.ascontiguousarray(q, np.float64).reshape(-1, 12)


def _i8(a, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.int8), (n,)))


def pcs_project(q, ac, ik, env=1):
    q = _q(q); n = len(q)
    d, s = _run("pcs_project", n, env, 0, q.tobytes() + _i8(ac, n).tobytes() + _i8(ik, n).tobytes())
    ok = np.frombuffer(d[:n], np.uint8).copy(); qo = np.frombuffer(d[n:], np.float64).reshape(n, 12).copy()
    return ok, qo, s


def pcs_validate(q, ac, ik, env=1):
    q = _q(q); n = len(q)
    d, s = _run("pcs_validate", n, env, 0, q.tobytes() + _i8(ac, n).tobytes() + _i8(ik, n).tobytes())
    ok = np.frombuffer(d[:n], np.uint8).copy(); valid = np.frombuffer(d[n:2 * n], np.uint8).copy()
    qo = np.frombuffer(d[2 * n:], np.float64).reshape(n, 12).copy()
    return ok, valid, qo, s


def spcs_seeded(seed, i0, n, env=1, timeout=None):
    d, s = _run("spcs_seeded", n, env, 0, struct.pack("<QQ", seed, i0), timeout=timeout)
    return np.frombuffer(d[:n], np.uint8).copy(), np.frombuffer(d[n:2 * n], np.uint8).copy(), s


def collide(q, env=1):
    q = _q(q); n = len(q)
    d, s = _run("collide", n, env, 0, q.tobytes())
    return np.frombuffer(d, np.int32).copy(), s


def rbs(qa, qb, ac, ik, env=1, count_nodes=True):
    qa = _q(qa); qb = _q(qb); n = len(qa)
    pay = np.concatenate([qa, qb], axis=1).tobytes() + _i8(ac, n).tobytes() + _i8(ik, n).tobytes()
    d, s = _run("rbs", n, env, 1 if count_nodes else 0, pay)
    return np.frombuffer(d[:n], np.uint8).copy(), np.frombuffer(d[n:], np.int32).copy(), s


def tridist(S, T, env=1):
    S = np.ascontiguousarray(S, np.float64).reshape(-1, 9); T = np.ascontiguousarray(T, np.float64).reshape(-1, 9); n = len(S)
    d, s = _run("tridist", n, env, 0, np.concatenate([S, T], axis=1).tobytes())
    a = np.frombuffer(d, np.float64)
    return a[:n].copy(), a[n:].reshape(n, 6).copy()


def pair_counters(q, env=1):
    """restated frames + pair table issued through the binary's own PQP_Tolerance; also the binary's collision_state"""
    q = _q(q); n = len(q)
    d, s = _run("pair_counters", n, env, 0, q.tobytes())
    flags = np.frombuffer(d[:n * MAX_PAIRS], np.uint8).reshape(n, MAX_PAIRS).copy()
    cnt = np.frombuffer(d[n * MAX_PAIRS:n * MAX_PAIRS + 8 * n], np.int32).reshape(n, 2).copy()
    ref = np.frombuffer(d[n * MAX_PAIRS + 8 * n:], np.int32).copy()
    return flags, cnt, ref


def min_distance(q, env=1):
    q = _q(q); n = len(q)
    d, s = _run("min_distance", n, env, 0, q.tobytes())
    return np.frombuffer(d[:8 * n], np.float64).copy(), np.frombuffer(d[8 * n:], np.int32).copy()


def identify_ik(q, env=1):
    q = _q(q); n = len(q)
    d, s = _run("identify_ik", n, env, 0, q.tobytes())
    return np.frombuffer(d, np.int32).reshape(n, 2).copy()


def isvalid_full(q, env=1):
    q = _q(q); n = len(q)
    d, s = _run("isvalid_full", n, env, 0, q.tobytes())
    return np.frombuffer(d[:n], np.uint8).copy()


def fk(q6, robot, env=1):
    q6 = np.ascontiguousarray(q6, np.float64).reshape(-1, 6); n = len(q6)
    d, s = _run("fk", n, env, 0, q6.tobytes() + _i8(robot, n).tobytes())
    return np.frombuffer(d, np.float64).reshape(n, 4, 4).copy()


def ik(T, robot, sol, env=1):
    T = np.ascontiguousarray(T, np.float64).reshape(-1, 16); n = len(T)
    d, s = _run("ik", n, env, 0, T.tobytes() + _i8(robot, n).tobytes() + _i8(sol, n).tobytes())
    return np.frombuffer(d[:n], np.uint8).copy(), np.frombuffer(d[n:], np.float64).reshape(n, 6).copy()


def reconstruct_rbs(qa, qb, ac, ik, env=1):
    """the binary's reconstructRBS(qa1, qa2, qb1, qb2, ac, ik, M = {a, b}, 0, 1, 1): (ok [n], list of row matrices)"""
    qa = _q(qa); qb = _q(qb); n = len(qa)
    pay = np.concatenate([qa, qb], axis=1).tobytes() + _i8(ac, n).tobytes() + _i8(ik, n).tobytes()
    d, s = _run("reconstruct_rbs", n, env, 0, pay)
    ok = np.frombuffer(d[:n], np.uint8).copy(); nrows = np.frombuffer(d[n:5 * n], np.int32).copy()
    rows = np.frombuffer(d[5 * n:], np.float64).reshape(-1, 12)
    out, o = [], 0
    for r in nrows:
        out.append(rows[o:o + r].copy()); o += r
    return ok, out


def ikproject5(q, ac, ik_nn, env=1):
    """the binary's IKproject(q1, q2, ik, active_chain, ik_nn): (valid, ac_out, ik [n, 2], q [n, 12])"""
    q = _q(q); n = len(q)
    nn = np.ascontiguousarray(ik_nn, np.int8).reshape(n, 2)
    d, s = _run("ikproject5", n, env, 0, q.tobytes() + _i8(ac, n).tobytes() + nn.tobytes())
    v = np.frombuffer(d[:n], np.uint8).copy(); aco = np.frombuffer(d[n:2 * n], np.int8).copy()
    iko = np.frombuffer(d[2 * n:10 * n], np.int32).reshape(n, 2).copy(); qo = np.frombuffer(d[10 * n:], np.float64).reshape(n, 12).copy()
    return v, aco, iko, qo
