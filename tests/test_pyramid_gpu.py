"""GPU parity for the pyramid driver (pyramidal_sgm.m): every level on the device through the C ABI vs the
numpy restatement whose per-level solver is the reference's own C++ (committed fixture) or the oracle run here.  Bit-exact."""
import os

import numpy as np
import pytest

from fsgm_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("H,W,n", [(37, 50, 2), (1, 7, 1), (7, 1, 1), (2, 2, 3), (375, 1242, 2), (94, 311, 1), (65, 33, 1)])
def test_impyramid_reduce(ctx, H, W, n):
    import torch
    from oracle import pyramid_oracle as pyo
    img = np.random.default_rng(H * 7 + W).integers(0, 256, (n, H, W)).astype(np.uint8)
    out = torch.empty((n, (H + 1) // 2, (W + 1) // 2), dtype=torch.uint8, device="cuda")
    ctx.impyramid_reduce_dev(_t(img), out)
    for i in range(n):
        assert np.array_equal(out.cpu().numpy()[i], pyo.reduce_closed_form(img[i])), i
    if H * W < 5000:
        assert np.array_equal(out.cpu().numpy()[0], pyo.reduce_general(img[0]))


def test_pyramid_driver_golden(ctx):
    from fsgm_b200 import api
    z = np.load(os.path.join(GOLD, "pyramid_a.npz"))
    L = int(z["numPyd"])
    o = api.pyd_opts(numPyd=L, ver=int(z["ver"]), hor=int(z["hor"]))
    mv, minC, lv = ctx.pyramidal_sgm(z["I0"], z["I1"], opts=o, levels=True)
    assert np.array_equal(minC, z["minC"])
    assert np.array_equal(mv, z["mv"])
    for i in range(L):
        assert np.array_equal(lv[i], z[f"mv_l{i}"]), i


@pytest.mark.parametrize("W,H,L,ver,hor,blocks", [(97, 61, 4, 2, 2, 1), (130, 70, 2, 3, 4, 3), (64, 48, 1, 2, 2, 2), (33, 35, 5, 1, 2, 1)])
def test_pyramid_driver_vs_oracle(ctx, oracle, W, H, L, ver, hor, blocks):
    """odd sizes (prior maps wider than the image), a single level, more levels than the image supports comfortably; batch of 2"""
    import torch
    from fsgm_b200 import api
    from oracle import pyramid_oracle as pyo
    solver = (lambda *a: oracle.ref_pyd(*a, stages=False)) if oracle.have_ref("pyd") else (lambda *a: oracle.port_pyd(*a, stages=False))
    ps = [synth.flow_pair(W, H, seed=W + i, umax=6, vmax=4, blocks=blocks) for i in range(2)]
    I0 = np.stack([p["I1"] for p in ps]); I1 = np.stack([p["I2"] for p in ps])
    o = api.pyd_opts(numPyd=L, ver=ver, hor=hor)
    ws, hs = api.pyramid_dims(W, H, L)
    mv = torch.empty((2, 2, H, W), dtype=torch.float64, device="cuda")
    minC = torch.empty((2, H, W), dtype=torch.int32, device="cuda")
    allv = torch.empty(sum(2 * 2 * w * h for w, h in zip(ws, hs)), dtype=torch.float64, device="cuda")
    ctx.pyramidal_sgm_dev(_t(I0), _t(I1), mv, minC, opts=o, mvPyd=allv)
    flat, off = allv.cpu().numpy(), 0
    levels = []
    for i in range(2):
        rmv, rmc, rlv = pyo.pyramidal_sgm(I0[i], I1[i], solver, numPyd=L, ver=ver, hor=hor)
        assert np.array_equal(minC.cpu().numpy().view(np.uint32)[i], rmc), i
        assert np.array_equal(mv.cpu().numpy()[i], rmv), i
        levels.append(rlv)
    for l, (w, h) in enumerate(zip(ws, hs)):
        blk = flat[off:off + 2 * 2 * w * h].reshape(2, 2, h, w)
        off += 2 * 2 * w * h
        for i in range(2):
            assert np.array_equal(blk[i], levels[i][l]), (l, i)


def test_pyramid_driver_matches_per_level_gateway_calls(ctx):
    """the on-device loop equals the MATLAB-style loop of host gateway calls (calc_pyd_cost_sgm per level) at half-KITTI size"""
    from fsgm_b200 import api
    from oracle import pyramid_oracle as pyo
    W, H, L = 621, 188, 3
    fp = synth.flow_pair(W, H, seed=3, umax=12, vmax=6)
    solver = lambda I1, I2, pre, rx, ry, agg, sub, P1, P2, diag, passes, adp: dict(zip(
        ("bestD", "minC", "mvSub"), ctx.calc_pyd_cost_sgm(I1, I2, pre, rx, ry, agg, sub, P1, P2, diag, passes, adp)))
    rmv, rmc, _ = pyo.pyramidal_sgm(fp["I1"], fp["I2"], solver, numPyd=L)
    mv, minC = ctx.pyramidal_sgm(fp["I1"], fp["I2"], opts=api.pyd_opts(numPyd=L))
    assert np.array_equal(mv, rmv) and np.array_equal(minC, rmc)


@pytest.mark.timeout(900)
def test_config_c_three_levels_full_kitti_size_vs_oracle(ctx, oracle):
    """BASELINE.json configs[2] at its real size: 3 levels of 1242x375 (-> 621x188 -> 311x94), r = 5 (121 labels), 8 paths,
    2 passes (pyramidal_sgm.m:14-22, :36-75), the on-device driver against the numpy restatement of the MATLAB loop whose
    per-level solver is the reference's own calc_pyd_cost_sgm build.  Level 0 alone is ~35 s of one host core."""
    from fsgm_b200 import api
    from oracle import pyramid_oracle as pyo
    solver = (lambda *a: oracle.ref_pyd(*a, stages=False)) if oracle.have_ref("pyd") else (lambda *a: oracle.port_pyd(*a, stages=False))
    W, H, L = 1242, 375, 3
    fp = synth.flow_pair(W, H, seed=1, umax=20, vmax=10)
    rmv, rmc, rlv = pyo.pyramidal_sgm(fp["I1"], fp["I2"], solver, numPyd=L)
    mv, minC, lv = ctx.pyramidal_sgm(fp["I1"], fp["I2"], opts=api.pyd_opts(numPyd=L), levels=True)
    assert np.array_equal(minC, rmc)
    for l in range(L):
        assert np.array_equal(lv[l], rlv[l]), l
    assert np.array_equal(mv, rmv)
