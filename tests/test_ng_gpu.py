"""GPU parity: neighbour-guided paths (calc_cost_sgm_ng, calc_pyd_cost_sgm_ng) vs golden vectors and the CPU oracle."""
import os

import numpy as np
import pytest

from fsgm_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _g(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def test_ng_vs_golden_with_stages(ctx):
    import torch
    g = _g("ng_a")
    H, W = g["I1"].shape
    minC, flow = ctx.calc_cost_sgm_ng(g["I1"], g["I2"], None, 1, 2, 0, int(g["P1"]), int(g["P2"]), seed=int(g["seed"]))
    assert np.array_equal(minC, g["minC"])
    assert np.array_equal(flow, g["flow"])
    # stage outputs through the device form
    dM = torch.empty((1, H, W), dtype=torch.int32, device="cuda")
    dF = torch.empty((1, 2, H, W), dtype=torch.float64, device="cuda")
    dSp = torch.empty((1, H, W, 108), dtype=torch.int32, device="cuda")
    dCe = torch.empty((1, H, W, 108, 3), dtype=torch.int32, device="cuda")
    ctx.calc_cost_sgm_ng_dev(_t(g["I1"][None]), _t(g["I2"][None]), int(g["P1"]), int(g["P2"]), dM, dF, seeds=[int(g["seed"])],
                             Sp=dSp, Centries=dCe)
    assert np.array_equal(dCe.cpu().numpy()[0], g["Centries"].astype(np.int32))
    assert np.array_equal(dSp.cpu().numpy()[0].view(np.uint32), g["Sp"].astype(np.uint32))


@pytest.mark.parametrize("W,H,P1,P2,seed", [(40, 28, 6, 32, 1), (33, 19, 6, 32, 7), (21, 30, 100, 250, 3), (2, 9, 6, 32, 1), (9, 1, 6, 32, 1)])
def test_ng_vs_oracle(ctx, oracle, W, H, P1, P2, seed):
    fp = synth.flow_pair(W, H, seed=W, umax=3, vmax=2)
    f = oracle.ref_ng if oracle.have_ref("ng") else oracle.port_ng
    want = f(fp["I1"], fp["I2"], P1, P2, seed=seed)
    minC, flow = ctx.calc_cost_sgm_ng(fp["I1"], fp["I2"], None, 1, 2, 0, P1, P2, seed=seed)
    assert np.array_equal(minC, want["minC"])
    assert np.array_equal(flow, want["flow"])


def test_ng_explicit_rand_stream_and_batch(ctx, oracle):
    """An explicit rand() stream gives the same result as the seed that generates it; pairs in a batch are independent."""
    import torch
    from fsgm_b200 import api
    W, H = 30, 22
    fp = [synth.flow_pair(W, H, seed=s, umax=3, vmax=2) for s in (1, 2, 3)]
    stream = api.glibc_rand(5, W * H * 8)
    a = ctx.calc_cost_sgm_ng(fp[0]["I1"], fp[0]["I2"], P1=6, P2=32, seed=5)
    b = ctx.calc_cost_sgm_ng(fp[0]["I1"], fp[0]["I2"], P1=6, P2=32, rand_stream=stream)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    I1 = _t(np.stack([p["I1"] for p in fp])); I2 = _t(np.stack([p["I2"] for p in fp]))
    dM = torch.empty((3, H, W), dtype=torch.int32, device="cuda")
    dF = torch.empty((3, 2, H, W), dtype=torch.float64, device="cuda")
    ctx.calc_cost_sgm_ng_dev(I1, I2, 6, 32, dM, dF, seeds=[5, 6, 7])
    for i, s in enumerate((5, 6, 7)):
        m, f = ctx.calc_cost_sgm_ng(fp[i]["I1"], fp[i]["I2"], P1=6, P2=32, seed=s)
        assert np.array_equal(dM.cpu().numpy()[i].view(np.uint32), m) and np.array_equal(dF.cpu().numpy()[i], f)


@pytest.mark.parametrize("name", ["pydng_a", "pydng_b"])
def test_pydng_vs_golden(ctx, name):
    import torch
    g = _g(name)
    H, W = g["I1"].shape
    r, agg, sub, P1, P2 = int(g["r"]), int(g["aggSize"]), int(g["sub"]), int(g["P1"]), int(g["P2"])
    minC, flow = ctx.calc_pyd_cost_sgm_ng(g["I1"], g["I2"], g["preMv"], r, agg, sub, P1, P2)
    assert np.array_equal(minC, g["minC"])
    assert np.array_equal(flow, g["flow"])
    S, D = 2 * r + 1, 9 * (2 * r + 1) ** 2
    dM = torch.empty((1, H, W), dtype=torch.int32, device="cuda")
    dF = torch.empty((1, 2, H, W), dtype=torch.float64, device="cuda")
    dSp = torch.empty((1, H, W, D), dtype=torch.int32, device="cuda")
    dC = torch.empty((1, H, W, D), dtype=torch.uint8, device="cuda")
    dXY = torch.empty((1, H, W, 2, 9, S), dtype=torch.int32, device="cuda")
    ctx.calc_pyd_cost_sgm_ng_dev(_t(g["I1"][None]), _t(g["I2"][None]), _t(g["preMv"][None]), r, agg, sub, P1, P2, dM, dF,
                                 Sp=dSp, cost=dC, XY=dXY)
    ce = g["Centries"].astype(np.int32).reshape(H, W, 9, S, S, 3)
    assert np.array_equal(dC.cpu().numpy()[0].reshape(H, W, 9, S, S), ce[..., 2])
    xy = dXY.cpu().numpy()[0]
    assert np.array_equal(np.broadcast_to(xy[:, :, 0, :, :, None], (H, W, 9, S, S)), ce[..., 0])
    assert np.array_equal(np.broadcast_to(xy[:, :, 1, :, None, :], (H, W, 9, S, S)), ce[..., 1])
    assert np.array_equal(dSp.cpu().numpy()[0].view(np.uint32), g["Sp"].astype(np.uint32))


@pytest.mark.parametrize("W,H,r,aggSize,sub,P1,P2,prior", [
    (36, 26, 1, 5, 1, 6, 32, "frac"), (30, 22, 2, 5, 0, 6, 32, "int"), (24, 18, 0, 3, 1, 6, 32, "frac"),
    (22, 16, 1, 5, 1, 100, 250, "frac"), (20, 14, 1, 1, 1, 6, 32, "wild"), (26, 15, 3, 5, 0, 6, 32, "int"),
    (28, 17, 2, 5, 1, 100, 250, "int"), (25, 14, 3, 5, 1, 6, 32, "frac"), (18, 12, 4, 5, 0, 6, 32, "int"), (16, 11, 5, 3, 1, 6, 32, "frac"),
    (31, 20, 1, 5, 0, 250, 100, "big")])
def test_pydng_vs_oracle(ctx, oracle, W, H, r, aggSize, sub, P1, P2, prior):
    fp = synth.flow_pair(W, H, seed=W + r, umax=2, vmax=2)
    rng = np.random.default_rng(W)
    mv = rng.normal(0, 2.0, (2, H - 3, W + 4))              # smaller than the image in y: hints are clamped to the map
    if prior == "int":
        mv = np.round(mv)
    if prior == "wild":
        mv[0, 1, 2], mv[1, 3, 4], mv[0, 5, 6] = np.nan, 1e300, -3e9
    if prior == "big":                                      # |mv| > 1 almost everywhere: regular candidate grids, fractional priors
        mv = mv + np.where(mv >= 0, 1.5, -1.5)
    f = oracle.ref_pydng if oracle.have_ref("pydng") else oracle.port_pydng
    want = f(fp["I1"], fp["I2"], mv, r, aggSize, sub, P1, P2)
    for generic in (0, 1):                                  # per-grid tables where the grids are regular / cell-by-cell search everywhere
        ctx.tune(7, generic)
        try:
            minC, flow = ctx.calc_pyd_cost_sgm_ng(fp["I1"], fp["I2"], mv, r, aggSize, sub, P1, P2)
        finally:
            ctx.tune(7, 0)
        assert np.array_equal(minC, want["minC"]), f"generic={generic}"
        assert np.array_equal(flow, want["flow"], equal_nan=True), f"generic={generic}"


@pytest.mark.timeout(600)
@pytest.mark.parametrize("W,H,seed", [(160, 64, 1), (131, 77, 4)])
def test_ng_larger_image_vs_oracle(ctx, oracle, W, H, seed):
    """calc_cost_sgm_ng beyond toy sizes (nothing above 40x28 was compared): ~10 k pixels, i.e. 80 k rand() draws and
    every ring-buffer wrap of the L1 / row hints exercised thousands of times."""
    fp = synth.flow_pair(W, H, seed=W + 1, umax=6, vmax=3)
    f = oracle.ref_ng if oracle.have_ref("ng") else oracle.port_ng
    want = f(fp["I1"], fp["I2"], 6, 32, seed=seed)
    minC, flow = ctx.calc_cost_sgm_ng(fp["I1"], fp["I2"], None, 1, 2, 0, 6, 32, seed=seed)
    assert np.array_equal(minC, want["minC"])
    assert np.array_equal(flow, want["flow"])


@pytest.mark.timeout(600)
@pytest.mark.parametrize("W,H,r,sub", [(160, 72, 1, 1), (128, 64, 2, 1)])
def test_pydng_larger_image_vs_oracle(ctx, oracle, W, H, r, sub):
    fp = synth.flow_pair(W, H, seed=W + r, umax=5, vmax=3)
    rng = np.random.default_rng(W + 7)
    mv = rng.normal(0, 2.0, (2, H, W))
    f = oracle.ref_pydng if oracle.have_ref("pydng") else oracle.port_pydng
    want = f(fp["I1"], fp["I2"], mv, r, 5, sub, 6, 32)
    minC, flow = ctx.calc_pyd_cost_sgm_ng(fp["I1"], fp["I2"], mv, r, 5, sub, 6, 32)
    assert np.array_equal(minC, want["minC"])
    assert np.array_equal(flow, want["flow"])
