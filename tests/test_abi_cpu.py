"""CPU-only checks of the product side: the C-ABI library loads and exports every symbol include/ckc.h declares, refuses to
compute without a GPU (no fallback), reports IO errors; host-side generator and sharding logic."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO


def _declared():
    src = open(os.path.join(REPO, "include", "ckc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ckc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from abb_dualarm_planning_b200 import load_library, lib_path
    L = load_library()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libckc.so does not export %s" % n
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path()]).decode()
    for n in names:
        assert re.search(r" T %s\b" % n, out), n


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from abb_dualarm_planning_b200 import Checker, CkcError
    with pytest.raises(CkcError) as e:
        Checker(env=1)
    assert "-3" in str(e.value)            # CKC_ERR_CUDA: the product path fails loudly, it never computes on the CPU


def test_bad_arguments_and_io_errors():
    from abb_dualarm_planning_b200 import load_library
    L = load_library()
    ctx = C.c_void_p()
    assert L.ckc_create(C.byref(ctx), 3, b"/nonexistent", 0) == -1          # CKC_ERR_ARG (env)
    assert L.ckc_create(C.byref(ctx), 1, b"/nonexistent/dir", 0) == -2      # CKC_ERR_IO, the reference exit(-1)s here
    assert not ctx.value
    assert L.ckc_pcs_validate(None, 1, None, None, None, None, None, None) < 0
    assert b"sm_100a" in L.ckc_version()


def test_oversized_mesh_is_refused_at_create(tmp_path):
    """the kernels pack (pair, node A, node B) and (pair, triangle A, triangle B) into one task word with 11 / 10-bit fields: a mesh with
    more than 1 024 triangles must be refused at ckc_create (CKC_ERR_ARG) instead of aliasing bits into the pair index; a bad pack and a
    truncated .tris file are I/O errors.  All of it is decided on the host before the first CUDA call, so this runs without a GPU."""
    import struct
    import numpy as np
    from abb_dualarm_planning_b200 import load_library
    from abb_dualarm_planning_b200.binding import MESH_PACK
    L = load_library()
    L.ckc_create_error.restype = C.c_char_p
    blob = open(MESH_PACK, "rb").read()
    assert blob[:8] == b"CKCMESH1"
    nm = struct.unpack("<i", blob[8:12])[0]
    off, out = 12, b"CKCMESH1" + struct.pack("<i", nm)
    rng = np.random.default_rng(0)
    for m in range(nm):
        mid, n = struct.unpack("<ii", blob[off:off + 8]); off += 8
        tri = blob[off:off + 72 * n]; off += 72 * n
        if m == 3:                                             # Link3_r.tris has 1 000 triangles: 25 more are one too many bits
            tri = tri + (rng.normal(size=(25, 9)) * 50).astype("<f8").tobytes(); n += 25
        out += struct.pack("<ii", mid, n) + tri
    big = tmp_path / "big.ckcmesh"; big.write_bytes(out)
    ctx = C.c_void_p()
    assert L.ckc_create(C.byref(ctx), 1, str(big).encode(), 0) == -1 and not ctx.value            # CKC_ERR_ARG
    assert b"1024 triangles" in L.ckc_create_error()
    bad = tmp_path / "bad.ckcmesh"; bad.write_bytes(blob[:1000])
    assert L.ckc_create(C.byref(ctx), 1, str(bad).encode(), 0) == -2 and not ctx.value            # CKC_ERR_IO
    d = tmp_path / "tris"; d.mkdir()
    (d / "Base_r.tris").write_text("3\n0 0 0 1 0 0 0 1 0\n")                                    # header says 3 triangles, one present
    assert L.ckc_create(C.byref(ctx), 1, str(d).encode(), 0) == -2 and not ctx.value


def test_product_does_not_touch_the_oracle():
    """a product path that routes through oracle/ would void every parity claim"""
    pkg = os.path.join(REPO, "abb_dualarm_planning_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "ckc_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt and '"oracle' not in txt, f
    out = subprocess.check_output(["ldd", os.path.join(pkg, "csrc", "libckc.so")]).decode()
    assert "oracle" not in out


def test_host_generator_matches_specification(oracle):
    from abb_dualarm_planning_b200 import sampling
    q, ik = sampling.sample_pcs(5, 123456, 3000)
    q2, ik2 = oracle.sample_pcs(5, 123456, 3000)
    assert (q == q2).all() and (ik == ik2).all()
    g = sampling.sample_gd(77, 2 ** 40, 500)
    assert (g == oracle.sample_gd(77, 2 ** 40, 500)).all()
    a = sampling.uniform(1, 0, 10, 0); b = sampling.uniform(2, 0, 10, 0)
    assert (a != b).any() and (0 <= a).all() and (a < 1).all()


def test_shard_ranges_partition():
    from abb_dualarm_planning_b200.sharding import shard_range
    for n in (0, 1, 7, 1000, 10 ** 8 + 3):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, %(repo)r)
import numpy as np, torch, torch.distributed as dist
from abb_dualarm_planning_b200 import sampling
from abb_dualarm_planning_b200.sharding import shard_range, reduce_scalar, gather_mask
from oracle import lib
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 6000
lo, hi = shard_range(n, rank, world)
orc = lib.Oracle(1)
q, ik = sampling.sample_pcs(21, lo, hi - lo)          # each rank generates ITS index range of the same stream
ok, valid, _ = orc.pcs_validate(q, 0, ik)             # (CPU stand-in for the per-rank GPU context: host logic under test)
mask = gather_mask(valid, dist)
t = reduce_scalar(1.0 + rank, "max", dist)
s = reduce_scalar(float(valid.sum()), "sum", dist)
if rank == 0:
    qa, ika = sampling.sample_pcs(21, 0, n)
    _, va, _ = orc.pcs_validate(qa, 0, ika)
    assert (mask == va).all(), "sharded mask differs from the single-rank mask"
    assert t == float(world) and s == float(va.sum())
    print("SHARD_OK", int(va.sum()))
dist.destroy_process_group()
'''


def test_two_rank_sharding_gloo(tmp_path):
    """world_size 2 on CPU (gloo): contiguous index ranges of the counter-generated stream, masks concatenated in rank order,
    max-over-ranks timing and summed counters -- no data-path collective."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"repo": REPO})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHARD_OK" in out.stdout


# ---- the drop-in C++ surface: every checker member the reference's planners call exists, flavour by flavour ----
FLAVOURS = ("PCS", "GD", "HB", "RX", "SG")


def _scanner():
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import scan_dropin_surface as S
    return S


def test_dropin_headers_declare_every_member_the_planners_call():
    """tests/golden/dropin_surface.json (tools/scan_dropin_surface.py over planners/*_{PCS,GD,HB,RX,SG}*.cpp and run/plan_*.cpp of the
    reference) lists every StateValidityChecker / two_robots / kdl / collisionDetection member those files call, with the argument
    counts used; the drop-in header of each flavour (run through the preprocessor) must declare an overload for each."""
    import json
    S = _scanner()
    surface = json.load(open(os.path.join(REPO, "tests", "golden", "dropin_surface.json")))
    assert sorted(surface) == sorted(FLAVOURS)
    assert sum(len(v) for v in surface.values()) >= 80
    for flav in FLAVOURS:
        assert S.missing(surface, flav) == [], "flavour %s" % flav
    if os.path.isdir(S.REF):                      # in the build container: the fixture is what a fresh scan produces
        fresh = S.scan_reference()
        assert {f: sorted(v) for f, v in fresh.items()} == {f: sorted(v) for f, v in surface.items()}


@pytest.mark.parametrize("flav", FLAVOURS)
def test_planner_style_subclass_compiles(flav):
    """host/surface_check.cpp: a class that inherits the drop-in checker publicly (as the reference's planners do) and calls each of
    those members with the reference's argument types -- syntax-checked by g++ for every flavour"""
    host = os.path.join(REPO, "abb_dualarm_planning_b200", "host")
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-DSURFACE_" + flav, os.path.join(host, "surface_check.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
