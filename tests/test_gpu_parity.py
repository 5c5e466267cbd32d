"""GPU parity: the CUDA path (through the C ABI of include/ckc.h) against the CPU oracle on the same seeded inputs and
against the golden vectors produced by the genuine reference code.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_Q = 1e-6          # rad: north-star tolerance for IK joint solutions (observed: ~1e-13)


@pytest.fixture(scope="module")
def ck():
    from abb_dualarm_planning_b200 import Checker
    c = Checker(env=1)
    yield c
    c.close()


def test_library_is_native(ck):
    import abb_dualarm_planning_b200 as pkg
    assert b"sm_100a" in ck.L.ckc_version()
    assert pkg.lib_path().endswith("csrc/libckc.so")
    assert ck.num_pairs == 116


def test_projection_vs_reference_binary_golden(ck, gold):
    g = gold("pcs_project.npz")
    for ac, okn, qn in ((0, "ok_ac0", "q_ac0"), (1, "ok_ac1", "q_ac1")):
        ok, q = ck.pcs_project(g["q"], ac, g["ik"])
        assert (ok == g[okn]).all()
        m = ok == 1
        assert np.abs(q[m] - g[qn][m]).max() < TOL_Q
        assert (q[~m] == g["q"][~m]).all()


def test_projection_vs_oracle_seeded(ck, oracle):
    q, ik = oracle.sample_pcs(7, 0, 50000)
    ok_o, q_o = oracle.pcs_project(q, 0, ik)
    ok, qo = ck.pcs_project(q, 0, ik)
    assert (ok == ok_o).all()
    assert np.abs(qo - q_o).max() < TOL_Q


def test_collision_per_pair_flags_vs_pqp_golden(ck, gold):
    g = gold("collide.npz")
    flags = np.unpackbits(g["pair_flags"], axis=1)[:, :116]
    got = ck.collide_pairs(g["q"])
    assert (got == flags).all()
    assert (ck.collide(g["q"]) == g["collision_state"]).all()


def test_reference_path_fixture_is_collision_free(ck, gold):
    g = gold("ref_path.npz")
    assert (ck.collide(g["path"]) == 0).all()


def test_validate_vs_oracle_seeded(ck, oracle):
    q, ik = oracle.sample_pcs(11, 0, 100000)
    ok_o, valid_o, q_o = oracle.pcs_validate(q, 0, ik)
    ok, valid, qo = ck.pcs_validate(q, 0, ik)
    assert (ok == ok_o).all() and (valid == valid_o).all()
    assert np.abs(qo - q_o).max() < TOL_Q


def test_million_sample_mask_vs_reference(ck, gold):
    """north star: mask identical to the reference's StateValidityCheckerPCS on the 1e6-sample seed set (generated on the
    device from the counter-based generator, checked against the masks the genuine reference binary produced)."""
    import torch
    g = gold("spcs_seed1_1e6_mask.npz")
    n = int(g["n"])
    ok_ref = np.unpackbits(g["ok_bits"])[:n]; valid_ref = np.unpackbits(g["valid_bits"])[:n]
    d_ok = torch.zeros(n, dtype=torch.uint8, device="cuda"); d_valid = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ck.spcs_validate_seeded_dev(1, 0, n, None, d_ok.data_ptr(), d_valid.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ok = d_ok.cpu().numpy(); valid = d_valid.cpu().numpy()
    assert (ok == ok_ref).all()
    assert (valid == valid_ref).all()
    assert valid.sum() == 11514


def test_identify_ik_and_limits(ck, oracle, gold):
    g = gold("identify_ik.npz")
    assert (ck.identify_ik(g["q"]) == g["ik"]).all()
    q, _ = oracle.sample_pcs(3, 0, 2000)
    q[::3] *= 1.3
    assert (ck.check_angle_limits(q) == oracle.check_angle_limits(q)).all()


def test_empty_and_tiny_batches(ck, oracle):
    ok, valid, qo = ck.pcs_validate(np.zeros((0, 12)), 0, np.zeros(0, np.int8))
    assert len(ok) == 0 and len(valid) == 0
    q, ik = oracle.sample_pcs(2, 0, 33)
    a = ck.pcs_validate(q[:1], 0, ik[:1]); b = oracle.pcs_validate(q[:1], 0, ik[:1])
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    a = ck.pcs_validate(q, 0, ik); b = oracle.pcs_validate(q, 0, ik)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
