"""GPU: size-independent properties at BASELINE.json's full sizes (no oracle needed at these sizes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ck():
    from abb_dualarm_planning_b200 import Checker
    c = Checker(env=1)
    yield c
    c.close()


def test_gd_million_host_and_device_paths_agree_and_are_idempotent(ck):
    """configs[1]: 1e6 S-GD samples.  (1) the device-resident SoA path, the pipelined host AoS path and the device-generated
    path give the same mask; (2) closure: every valid configuration passes the RX residual test at 1e-3; (3) idempotence:
    re-validating the projected valid configurations keeps them valid and moves them by < 2e-4 rad (the first result carries the
    solver's 1e-5 stopping residual plus the 7e-6 rad perturbation of wrapping with 2*3.1416; one more Newton step removes it)."""
    import torch
    from abb_dualarm_planning_b200 import sampling
    n = 1_000_000
    q = sampling.sample_gd(2024, 0, n)
    valid_h, qo_h = ck.gd_validate(q)
    dq = torch.from_numpy(q).cuda().t().contiguous()
    dqo = torch.empty((12, n), dtype=torch.float64, device="cuda"); dv = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ck.gd_validate_dev(n, dq.data_ptr(), dqo.data_ptr(), dv.data_ptr(), st)
    torch.cuda.synchronize()
    assert (dv.cpu().numpy() == valid_h).all()
    dv2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    ck.sgd_validate_seeded_dev(2024, 0, n, None, None, dv2.data_ptr(), st)
    torch.cuda.synchronize()
    assert (dv2.cpu().numpy() == valid_h).all()
    m = valid_h == 1
    assert 10000 < m.sum() < 20000                       # ~1.4 % of uniform seeds end valid
    assert np.abs(dqo.t().cpu().numpy()[m] - qo_h[m]).max() == 0.0
    v_rx, res = ck.rx_validate(qo_h[m], 1e-3)
    assert v_rx.all() and res.max() < 1e-3               # closed chain (residual mixes 0.003*mm and rad), limits, collision free
    valid2, qo2 = ck.gd_validate(qo_h[m])
    assert valid2.all()
    d = np.abs(qo2 - qo_h[m]); d = np.minimum(d, np.abs(d - 2 * 3.1416))
    assert d.max() < 2e-4


def test_pcs_ten_million_seeded_checksum(ck):
    """S-PCS, 1e7 device-generated samples: the first 1e6 reproduce the reference mask (golden), the rest must keep the same
    statistics; the chunking / fan-out over streams must not change a single flag (two different chunkings compared)."""
    import os
    import torch
    from conftest import GOLD
    n = 10_000_000
    st = torch.cuda.current_stream().cuda_stream
    ok = torch.empty(n, dtype=torch.uint8, device="cuda"); valid = torch.empty(n, dtype=torch.uint8, device="cuda")
    ck.spcs_validate_seeded_dev(1, 0, n, None, ok.data_ptr(), valid.data_ptr(), st)
    torch.cuda.synchronize()
    g = np.load(os.path.join(GOLD, "spcs_seed1_1e6_mask.npz"))
    assert (valid[:1_000_000].cpu().numpy() == np.unpackbits(g["valid_bits"])[:1_000_000]).all()
    ck.set_overlap(False)
    ok2 = torch.empty(n, dtype=torch.uint8, device="cuda"); valid2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    ck.spcs_validate_seeded_dev(1, 0, n, None, ok2.data_ptr(), valid2.data_ptr(), st)
    torch.cuda.synchronize()
    ck.set_overlap(True)
    assert torch.equal(ok, ok2) and torch.equal(valid, valid2)
    f_ok, f_valid = ok.float().mean().item(), valid.float().mean().item()
    assert 0.0230 < f_ok < 0.0245 and 0.0110 < f_valid < 0.0120


def test_full_size_1e8_seeded_sweeps(ck):
    """BASELINE configs[4] at its full size on one GPU: 1e8 S-PCS and 1e8 S-GD samples generated on the device.  The first 1e6
    flags reproduce the golden masks, one 1e8 call equals ten 1e7 calls flag for flag (sample i depends on (seed, i) only, so this
    is also what sharding the index range over ranks returns), and the feasibility statistics stay in the reference's band."""
    import os
    import torch
    from conftest import GOLD
    n, part = 100_000_000, 10_000_000
    st = torch.cuda.current_stream().cuda_stream
    for name, call, gold_file, band in (
            ("pcs", lambda i0, m, v: ck.spcs_validate_seeded_dev(1, i0, m, None, None, v, st), "spcs_seed1_1e6_mask.npz", (0.0110, 0.0120)),
            ("gd", lambda i0, m, v: ck.sgd_validate_seeded_dev(1, i0, m, None, None, v, st), "sgd_seed1_1e6_mask.npz", (0.0135, 0.0147))):
        whole = torch.empty(n, dtype=torch.uint8, device="cuda"); parts = torch.empty(n, dtype=torch.uint8, device="cuda")
        call(0, n, whole.data_ptr())
        for i0 in range(0, n, part):
            call(i0, part, parts[i0:].data_ptr())
        torch.cuda.synchronize()
        g = np.load(os.path.join(GOLD, gold_file))
        want = np.unpackbits(g["valid_bits"])[:1_000_000]
        got = whole[:1_000_000].cpu().numpy()
        assert (got != want).sum() <= (2 if name == "gd" else 0), name
        assert torch.equal(whole, parts), name
        f = whole.float().mean().item()
        assert band[0] < f < band[1], (name, f)
        del whole, parts


def test_rbs_is_symmetric_and_monotone(ck, gold):
    """checkMotionRBS(a, b) == checkMotionRBS(b, a) (the bisection tree is mirrored), and an edge that succeeds keeps
    succeeding when cut in half at its (projected) midpoint."""
    g = gold("rbs.npz")
    ok_ab, nc_ab = ck.pcs_check_motion_rbs(g["qa"], g["qb"], 0, g["ik"])
    ok_ba, nc_ba = ck.pcs_check_motion_rbs(g["qb"], g["qa"], 0, g["ik"])
    assert (ok_ab == ok_ba).all() and (nc_ab[ok_ab == 1] == nc_ba[ok_ab == 1]).all()
    good = ok_ab == 1
    mid = (g["qa"][good] + g["qb"][good]) / 2
    okm, vm, qm = ck.pcs_validate(mid, 0, g["ik"][good])
    assert vm.all()
    ok_l, _ = ck.pcs_check_motion_rbs(g["qa"][good], qm, 0, g["ik"][good])
    ok_r, _ = ck.pcs_check_motion_rbs(qm, g["qb"][good], 0, g["ik"][good])
    assert ok_l.all() and ok_r.all()


def test_pcs_four_million_mask_vs_oracle(ck, oracle):
    """a second, larger seed set (seed 77, 4e6 S-PCS samples, ~95 k collision checks): the device-generated masks against the
    CPU oracle -- any FP32 box-culling error inside the 8 mm band or any traversal bug would show up as a flipped flag."""
    import torch
    from abb_dualarm_planning_b200 import sampling
    n = 4_000_000
    st = torch.cuda.current_stream().cuda_stream
    ok = torch.empty(n, dtype=torch.uint8, device="cuda"); valid = torch.empty(n, dtype=torch.uint8, device="cuda")
    ck.spcs_validate_seeded_dev(77, 0, n, None, ok.data_ptr(), valid.data_ptr(), st)
    torch.cuda.synchronize()
    ok = ok.cpu().numpy(); valid = valid.cpu().numpy()
    for i0 in range(0, n, 500_000):
        q, ik = sampling.sample_pcs(77, i0, 500_000)
        ok_o, valid_o, _ = oracle.pcs_validate(q, 0, ik)
        assert (ok[i0:i0 + 500_000] == ok_o).all()
        assert (valid[i0:i0 + 500_000] == valid_o).all()
    assert 90_000 < ok.sum() < 100_000
