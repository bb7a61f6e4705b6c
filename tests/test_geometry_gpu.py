"""GPU parity for the dense epipolar prologue / epilogue through the C ABI vs oracle/geometry_oracle.py.
fp64 maps are compared bit for bit; the fused call is compared with prologue -> gateway 1 -> epilogue done separately."""
import numpy as np
import pytest

from fsgm_b200 import synth

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("W,H,n,rot", [(120, 80, 3, 0.4), (1242, 375, 2, 0.2), (33, 7, 1, 3.0), (1, 1, 2, 0.1)])
def test_geometry_maps_bit_exact(ctx, W, H, n, rot):
    import torch
    from oracle import geometry_oracle as go
    cams = [synth.epipolar_camera(W, H, seed=10 + i, rot_deg=rot) for i in range(n)]
    Pd0 = torch.empty((n, 2, H, W), dtype=torch.float64, device="cuda"); dirn = torch.empty_like(Pd0); Rf = torch.empty_like(Pd0)
    O = torch.empty((n, H, W), dtype=torch.float64, device="cuda")
    ctx.epipolar_geometry_dev([c["F"] for c in cams], [c["H"] for c in cams], [c["epi"] for c in cams],
                              [c["direction"] for c in cams], Pd0, dirn, O, Rf)
    for i, c in enumerate(cams):
        r = go.epipolar_geometry(c["F"], c["H"], c["epi"], c["direction"], W, H)
        for got, want, name in zip((Pd0, dirn, O, Rf), r, ("Pd0", "dirn", "O", "Rflow")):
            assert np.array_equal(got[i].cpu().numpy(), want), (i, name)
    # epilogue
    best = np.random.default_rng(W).integers(0, 60000, (n, H, W)).astype(np.uint32)
    flow = torch.empty_like(Pd0)
    ctx.epipolar_flow_dev(_t(best.view(np.int32)), dirn, Rf, flow)
    for i in range(n):
        assert np.array_equal(flow[i].cpu().numpy(), go.epipolar_flow(best[i], dirn[i].cpu().numpy(), Rf[i].cpu().numpy())), i


def test_geometry_degenerate_inputs(ctx):
    """zero F (unnormalised line), singular H (division by zero -> inf/NaN): same values as numpy (NaN positions equal)"""
    import torch
    from oracle import geometry_oracle as go
    W, H = 16, 9
    F = np.zeros(9); F[8] = 1.0
    Hs = np.array([1.0, 0, 0, 0, 1, 0, 0.1, 0, -0.5])                   # third row vanishes at x = 5
    Pd0 = torch.empty((1, 2, H, W), dtype=torch.float64, device="cuda"); dirn = torch.empty_like(Pd0); Rf = torch.empty_like(Pd0)
    O = torch.empty((1, H, W), dtype=torch.float64, device="cuda")
    ctx.epipolar_geometry_dev(F, Hs, [3.0, 2.0], [1], Pd0, dirn, O, Rf)
    r = go.epipolar_geometry(F, Hs, [3.0, 2.0], 1, W, H)
    for got, want in zip((Pd0, dirn, O, Rf), r):
        assert np.array_equal(got[0].cpu().numpy(), want, equal_nan=True)
    assert not np.isfinite(r[0]).all()


@pytest.mark.parametrize("W,H,D,n", [(160, 96, 64, 2), (200, 64, 256, 17)])
def test_epipolar_sgm_of_fused_call(ctx, oracle, W, H, D, n):
    """host images + F/H/epipole -> flow: equals oracle geometry -> gateway 1 -> oracle epilogue; first pair also vs the CPU oracle"""
    from fsgm_b200 import api
    from oracle import geometry_oracle as go
    o = api.epi_opts(paths=8)
    ps = [synth.epipolar_pair(W, H, D, seed=50 + i) for i in range(n)]
    cams = [synth.epipolar_camera(W, H, seed=60 + i, rot_deg=0.05) for i in range(n)]
    I0 = np.stack([p["I1"] for p in ps]); I1 = np.stack([p["I2"] for p in ps])
    flow, minC = ctx.epipolar_sgm_of_batch(I0, I1, [c["F"] for c in cams], [c["H"] for c in cams], [c["epi"] for c in cams],
                                           [c["direction"] for c in cams], D, 0.3, 6, 64, opts=o)
    for i in (0, n - 1):
        c = cams[i]
        Pd0, dirn, O, Rf = go.epipolar_geometry(c["F"], c["H"], c["epi"], c["direction"], W, H)
        bestD, mC, _, _ = ctx.calc_cost_sgm(I0[i], I1[i], D, 0.3, Pd0, dirn, O, 6, 64, opts=o)
        assert np.array_equal(minC[i], mC), i
        assert np.array_equal(flow[i], go.epipolar_flow(bestD, dirn, Rf)), i
        if i == 0:
            f = oracle.ref_epi if oracle.have_ref("epi8") else oracle.port_epi
            ref = f(I0[i], I1[i], D, 0.3, Pd0, dirn, O, 6, 64, paths=8, stages=False)
            assert np.array_equal(minC[i], ref["minC"])
            want = go.epipolar_flow(ref["bestD"], dirn, Rf)
            assert np.array_equal(flow[i].reshape(2, -1)[:, :-1], want.reshape(2, -1)[:, :-1])


def test_epipolar_sgm_of_float_form_is_the_rounded_fp64_flow(ctx):
    """The CV_32FC2 form (float, u/v interleaved; proj/include/epi_sgm.h:6-12) is the fp64 flow rounded to nearest, with and
    without the optional minC, through the same enqueue-only pipeline."""
    from fsgm_b200 import api
    W, H, D, n = 150, 64, 64, 3
    o = api.epi_opts(paths=8)
    ps = [synth.epipolar_pair(W, H, D, seed=70 + i) for i in range(n)]
    cams = [synth.epipolar_camera(W, H, seed=80 + i, rot_deg=0.05) for i in range(n)]
    I0 = np.stack([p["I1"] for p in ps]); I1 = np.stack([p["I2"] for p in ps])
    g = ([c["F"] for c in cams], [c["H"] for c in cams], [c["epi"] for c in cams], [c["direction"] for c in cams])
    flow, minC = ctx.epipolar_sgm_of_batch(I0, I1, *g, D, 0.3, 6, 64, opts=o)
    f32, m32 = ctx.epipolar_sgm_of_batch(I0, I1, *g, D, 0.3, 6, 64, opts=o, f32=True)
    assert f32.shape == (n, H, W, 2) and f32.dtype == np.float32
    assert np.array_equal(m32, minC)
    assert np.array_equal(f32, np.moveaxis(flow, 1, -1).astype(np.float32))
    only = np.zeros((n, H, W, 2), np.float32)
    ctx.epipolar_sgm_of_batch(I0, I1, *g, D, 0.3, 6, 64, opts=o, f32=True, out=(only, None))
    assert np.array_equal(only, f32)
