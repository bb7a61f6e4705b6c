"""GPU parity: pyramidal 2-D-window path (calc_pyd_cost_sgm) through the C ABI vs golden vectors and the CPU oracle."""
import os

import numpy as np
import pytest

from fsgm_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _oracle_pyd(oracle, *a, **k):
    return (oracle.ref_pyd if oracle.have_ref("pyd") else oracle.port_pyd)(*a, **k)


@pytest.mark.parametrize("name", ["pyd_a", "pyd_b", "pyd_c"])
def test_pyd_vs_golden(ctx, name):
    import torch
    z = np.load(os.path.join(GOLD, name + ".npz"))
    g = {k: z[k] for k in z.files}
    rx, ry, agg, sub = int(g["rx"]), int(g["ry"]), int(g["agg"]), int(g["sub"])
    P1, P2, diag, passes, adp = int(g["P1"]), int(g["P2"]), int(g["diag"]), int(g["passes"]), int(g["adaptive"])
    H, W = g["I1"].shape
    D = (2 * rx + 1) * (2 * ry + 1)
    # whole gateway, host arrays
    bestD, minC, mvSub = ctx.calc_pyd_cost_sgm(g["I1"], g["I2"], g["preMv"], rx, ry, agg, sub, P1, P2, diag, passes, adp)
    assert np.array_equal(minC, g["minC"])
    assert np.array_equal(bestD, g["bestD"])
    assert np.array_equal(mvSub, g["mvSub"])          # fp64 division is IEEE on both sides: expect bit equality
    # stages: cost volume from our census, Sp from the golden C
    I1, I2, mv = _t(g["I1"][None]), _t(g["I2"][None]), _t(g["preMv"][None])
    cen1 = torch.empty((1, H, W), dtype=torch.int32, device="cuda")
    cen2 = torch.empty_like(cen1)
    ctx.census_dev(I1, cen1)
    ctx.census_dev(I2, cen2)
    Cv = torch.empty((1, H, W, D), dtype=torch.uint8, device="cuda")
    ctx.pyd_cost_dev(cen1, cen2, mv, agg, rx, ry, Cv)
    assert np.array_equal(Cv.cpu().numpy()[0], g["C"])
    Sp = torch.empty((1, H, W, D), dtype=torch.int16, device="cuda")
    b = torch.empty((1, H, W), dtype=torch.int32, device="cuda")
    m = torch.empty_like(b)
    s = torch.empty((1, 2, H, W), dtype=torch.float64, device="cuda")
    ctx.pyd_aggregate_dev(_t(g["C"][None]), I1, mv, rx, ry, sub, P1, P2, diag, passes, adp, b, m, s, Sp=Sp)
    assert np.array_equal(Sp.cpu().numpy().view(np.uint16)[0], g["Sp"])


CASES = [  # W, H, rx, ry, agg, sub, P1, P2, diag, passes, adaptive, prior
    (60, 40, 5, 5, 2, 1, 6, 32, 1, 2, 0, "zero"),        # reference defaults (pyramidal_sgm.m:14-22), D = 121
    (60, 40, 4, 4, 2, 1, 6, 32, 1, 2, 0, "int"),         # BASELINE's +-4 window, D = 81
    (45, 33, 3, 2, 2, 1, 6, 32, 1, 2, 1, "frac"),        # fractional prior, adaptive P2, D = 35
    (45, 33, 2, 6, 1, 0, 6, 32, 0, 2, 0, "int"),         # no diagonals, 3x3 aggregation
    (40, 30, 7, 8, 2, 1, 6, 32, 1, 1, 0, "int"),         # single pass, D = 255 (NJ = 8)
    (32, 24, 1, 1, 0, 1, 200, 250, 1, 2, 0, "frac"),     # mod-256 domain, no aggregation
    (30, 20, 2, 2, 2, 1, 6, 32, 1, 3, 0, "int"),         # totalPass = 3 repeats the reversed sweeps
    (30, 20, 2, 2, 2, 1, 6, 32, 1, 0, 0, "int"),         # totalPass = 0: Sp stays zero
    (24, 18, 0, 0, 2, 1, 6, 32, 1, 2, 0, "wild"),        # single label; NaN / huge prior values
]


@pytest.mark.parametrize("W,H,rx,ry,agg,sub,P1,P2,diag,passes,adp,prior", CASES)
def test_pyd_vs_oracle(ctx, oracle, W, H, rx, ry, agg, sub, P1, P2, diag, passes, adp, prior):
    fp = synth.flow_pair(W, H, seed=W + rx, umax=max(1, rx - 1), vmax=max(1, ry - 1))
    rng = np.random.default_rng(W * 7 + H)
    mv = np.zeros((2, H + 3, W + 5))
    if prior == "int":
        mv[:] = rng.integers(-3, 4, mv.shape)
    elif prior == "frac":
        mv[:] = rng.normal(0, 2.0, mv.shape)
    elif prior == "wild":
        mv[:] = rng.normal(0, 2.0, mv.shape)
        mv[0, 2, 3], mv[1, 4, 5], mv[0, 6, 7], mv[1, 1, 1] = np.nan, 1e300, -1e300, 3e9
    want = _oracle_pyd(oracle, fp["I1"], fp["I2"], mv, rx, ry, agg, sub, P1, P2, diag, passes, adp)
    bestD, minC, mvSub = ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv, rx, ry, agg, sub, P1, P2, diag, passes, adp)
    assert np.array_equal(minC, want["minC"])
    assert np.array_equal(bestD, want["bestD"])
    assert np.array_equal(mvSub, want["mvSub"], equal_nan=True)
    assert np.nanmax(np.abs(mvSub - want["mvSub"])) <= 1e-4 if mvSub.size else True     # the contract's tolerance


def test_pyd_sweep_each_direction(ctx, oracle):
    import ctypes as C
    import torch
    rng = np.random.default_rng(8)
    H, W, rx, ry = 21, 29, 3, 2
    Sx, Sy = 2 * rx + 1, 2 * ry + 1
    D = Sx * Sy
    Cv = rng.integers(0, 26, (H, W, D), dtype=np.uint8)
    I1 = rng.integers(0, 256, (H, W), dtype=np.uint8)
    mv = rng.normal(0, 1.5, (2, H, W + 2))
    lib = oracle._port()
    for r in range(8):
        want = np.empty((H, W, D), np.uint8)
        lib.orc_sweep2d(Cv.ctypes.data_as(C.c_void_p), I1.ctypes.data_as(C.c_void_p), W, H, mv.ctypes.data_as(C.c_void_p),
                        W + 2, H, Sx, Sy, 6, 32, 1, r, want.ctypes.data_as(C.c_void_p))
        L = torch.empty((1, H, W, D), dtype=torch.uint8, device="cuda")
        ctx.pyd_sweep_dev(_t(Cv[None]), _t(I1[None]), _t(mv[None]), rx, ry, 6, 32, 1, r, L)
        assert np.array_equal(L.cpu().numpy()[0], want), r


def test_pyd_bad_arguments(ctx):
    from fsgm_b200 import api
    fp = synth.flow_pair(16, 12, seed=1)
    with pytest.raises(api.FsgmError):          # prior smaller than the image
        ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], np.zeros((2, 12, 15)), 1, 1, 2, 1, 6, 32)
    with pytest.raises(api.FsgmError) as e:     # 33*33 labels
        ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], np.zeros((2, 12, 16)), 16, 16, 2, 1, 6, 32)
    assert e.value.code == api.FSGM_ERR_DOMAIN


def test_pyd_half_kitti_size_vs_oracle(ctx, oracle):
    """Config C's middle pyramid level at its real size (621x188, r = 5 -> 121 labels, 8 paths, 2 passes): the whole
    gateway against the CPU oracle (about 10 s of CPU with the reference build)."""
    W, H = 621, 188
    fp = synth.flow_pair(W, H, seed=21, umax=4, vmax=3)
    rng = np.random.default_rng(2)
    mv = np.round(rng.normal(0, 1.5, (2, H, W)))
    want = _oracle_pyd(oracle, fp["I1"], fp["I2"], mv, 5, 5, 2, 1, 6, 32, 1, 2, 0, stages=False)
    bestD, minC, mvSub = ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv, 5, 5, 2, 1, 6, 32, 1, 2, 0)
    assert np.array_equal(minC, want["minC"])
    assert np.array_equal(bestD, want["bestD"])
    assert np.array_equal(mvSub, want["mvSub"])


def _prior(kind, rng, H, W, pad=(3, 5)):
    mv = np.zeros((2, H + pad[0], W + pad[1]))
    if kind == "blocks":          # what the pyramid driver produces: 2 x integer flow, constant on 2x2 blocks (mostly zero shifts)
        coarse = rng.integers(-2, 3, (2, (H + pad[0] + 1) // 2, (W + pad[1] + 1) // 2))
        mv[:] = 2.0 * np.repeat(np.repeat(coarse, 2, axis=1), 2, axis=2)[:, :mv.shape[1], :mv.shape[2]]
    elif kind == "int":
        mv[:] = rng.integers(-4, 5, mv.shape)
    elif kind == "frac":
        mv[:] = rng.normal(0, 2.0, mv.shape)
    elif kind == "far":           # shifts far beyond the window in both signs, exact .5 ties, NaN / inf / huge
        mv[:] = rng.integers(-40, 41, mv.shape) * 0.5
        mv[0, 2, 3], mv[1, 4, 5], mv[0, 6, 7], mv[1, 1, 1], mv[0, 5, 1] = np.nan, 1e300, -1e300, 3e9, np.inf
    return mv


# (W, H, rx, ry, agg, sub, P1, P2, prior kind, forced cluster size): the row-synchronous cluster kernels of pydv.cu at every
# cluster shape they can take (1 = one CTA per pair, non-portable sizes above 8), window shapes up to 11 x 11, strips of 2 columns
CLUSTER_CASES = [
    (60, 40, 5, 5, 2, 1, 6, 32, "zero", 1), (60, 40, 5, 5, 2, 1, 6, 32, "blocks", 2), (61, 33, 5, 5, 2, 1, 6, 32, "int", 3),
    (75, 28, 4, 4, 2, 1, 6, 32, "frac", 4), (90, 24, 5, 4, 1, 0, 6, 32, "blocks", 9), (64, 20, 3, 5, 2, 1, 6, 32, "far", 16),
    (48, 27, 2, 1, 2, 1, 10, 60, "frac", 8), (40, 30, 0, 5, 1, 1, 6, 32, "int", 5), (40, 30, 5, 0, 2, 1, 0, 0, "int", 2),
    (21, 50, 1, 1, 2, 1, 6, 32, "far", 7), (130, 21, 5, 5, 2, 1, 6, 32, "blocks", 0), (47, 47, 4, 5, 2, 1, 6, 32, "far", 0),
]


@pytest.mark.parametrize("W,H,rx,ry,agg,sub,P1,P2,prior,cs", CLUSTER_CASES)
def test_pyd_cluster_path_vs_oracle(ctx, oracle, W, H, rx, ry, agg, sub, P1, P2, prior, cs):
    """The cluster path (padded-grid volumes, three directions per pass in shared memory, WTA fused) against the reference build,
    and against the one-warp-per-scanline kernels it replaces (fsgm_tune key 5 = -1)."""
    fp = synth.flow_pair(W, H, seed=W + rx, umax=max(1, rx - 1), vmax=max(1, ry - 1))
    mv = _prior(prior, np.random.default_rng(W * 3 + H + cs), H, W)
    want = _oracle_pyd(oracle, fp["I1"], fp["I2"], mv, rx, ry, agg, sub, P1, P2, 1, 2, 0)
    try:
        ctx.tune(5, cs)
        ctx.profile(True); ctx.profile_reset()
        bestD, minC, mvSub = ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv, rx, ry, agg, sub, P1, P2, 1, 2, 0)
        st = ctx.profile_read(); ctx.profile(False)
        if cs > 0:
            assert st["pyd_sweep"][1] == 4, st              # shift flags, horizontal pair, down pass, up pass: the cluster path ran
        else:
            assert st["pyd_sweep"][1] == 2, st              # shift descriptors + one launch for all directions: lane = path kernels
        assert "pyd_wta" in st and st["pyd_wta"][1] == 1
        ctx.tune(5, -1)
        b2, m2, s2 = ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv, rx, ry, agg, sub, P1, P2, 1, 2, 0)
    finally:
        ctx.profile(False)
        ctx.tune(5, 0)
    assert np.array_equal(minC, want["minC"])
    assert np.array_equal(bestD, want["bestD"])
    assert np.array_equal(mvSub, want["mvSub"], equal_nan=True)
    assert np.array_equal(minC, m2) and np.array_equal(bestD, b2) and np.array_equal(mvSub, s2, equal_nan=True)


# (W, H, rx, ry, agg, sub, P1, P2, diag, passes, adaptive, prior kind): the lane = path kernels of pydl.cu (one thread per scanline
# and direction, volumes as [y][label column][x][16-byte frame]) on every window width they are instantiated for, with
# truncation-toward-zero duplicates in both axes (negative integer shifts), shifts beyond the window, NaN / inf / huge priors,
# exact .5 ties, adaptive P2, no diagonals, a single pass, images narrower than a warp and wider than two
LANE_CASES = [
    (60, 40, 5, 5, 2, 1, 6, 32, 1, 2, 0, "zero"), (70, 37, 5, 5, 2, 1, 6, 32, 1, 2, 0, "blocks"), (61, 33, 5, 5, 2, 1, 6, 32, 1, 2, 0, "int"),
    (75, 28, 4, 4, 2, 1, 6, 32, 1, 2, 0, "frac"), (90, 24, 5, 4, 1, 0, 6, 32, 1, 2, 1, "blocks"), (64, 20, 3, 5, 2, 1, 6, 32, 1, 2, 0, "far"),
    (48, 27, 2, 1, 2, 1, 10, 60, 0, 2, 0, "frac"), (40, 30, 0, 5, 1, 1, 6, 32, 1, 1, 0, "int"), (40, 30, 5, 0, 2, 1, 0, 0, 1, 2, 1, "int"),
    (21, 50, 1, 1, 2, 1, 6, 32, 1, 2, 0, "far"), (130, 21, 5, 5, 2, 1, 6, 32, 1, 2, 1, "int"), (47, 47, 4, 5, 2, 1, 100, 105, 1, 2, 0, "far"),
    (5, 70, 5, 5, 2, 1, 6, 32, 1, 2, 0, "int"), (33, 3, 3, 3, 2, 1, 6, 32, 1, 2, 0, "frac"),
]


@pytest.mark.parametrize("forced", [0, -2])
@pytest.mark.parametrize("W,H,rx,ry,agg,sub,P1,P2,diag,passes,adp,prior", LANE_CASES)
def test_pyd_lane_path_vs_oracle(ctx, oracle, W, H, rx, ry, agg, sub, P1, P2, diag, passes, adp, prior, forced):
    """forced = -2 sends every shifted step through the label-by-label form that the kernel keeps for prior differences whose
    (int)(s + dd + 0.5) map is not `shift, plus one below zero` (fsgm_tune key 5)."""
    fp = synth.flow_pair(W, H, seed=W + rx, umax=max(1, rx - 1), vmax=max(1, ry - 1))
    mv = _prior(prior, np.random.default_rng(W * 3 + H), H, W)
    want = _oracle_pyd(oracle, fp["I1"], fp["I2"], mv, rx, ry, agg, sub, P1, P2, diag, passes, adp)
    try:
        ctx.tune(5, forced)
        ctx.profile(True); ctx.profile_reset()
        bestD, minC, mvSub = ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv, rx, ry, agg, sub, P1, P2, diag, passes, adp)
        st = ctx.profile_read()
    finally:
        ctx.profile(False)
        ctx.tune(5, 0)
    assert st["pyd_sweep"][1] == 2, st
    assert np.array_equal(minC, want["minC"])
    assert np.array_equal(bestD, want["bestD"])
    assert np.array_equal(mvSub, want["mvSub"], equal_nan=True)


def test_pyd_lane_path_rounding_anomaly(ctx, oracle):
    """Prior differences just below .5: (double)s + dd rounds up for some label rows only, so (int)(s + dd + 0.5) is not a
    shift — the descriptor pre-pass must flag these steps (the result then comes from the label-by-label form)."""
    W, H = 40, 24
    fp = synth.flow_pair(W, H, seed=5, umax=2, vmax=2)
    mv = np.zeros((2, H, W))
    eps = 2.0 ** -53
    mv[0, :, 1::2] = 0.5 - eps
    mv[1, 1::2, :] = -(0.5 - eps)
    mv[0, 5:9, 7:20] += 1.0
    want = _oracle_pyd(oracle, fp["I1"], fp["I2"], mv, 5, 5, 2, 1, 6, 32, 1, 2, 0)
    bestD, minC, mvSub = ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv, 5, 5, 2, 1, 6, 32, 1, 2, 0)
    assert np.array_equal(minC, want["minC"])
    assert np.array_equal(bestD, want["bestD"])
    assert np.array_equal(mvSub, want["mvSub"], equal_nan=True)


def test_pyd_cluster_batch_pairs_are_independent(ctx, oracle):
    import torch
    W, H, n = 70, 26, 5
    fps = [synth.flow_pair(W, H, seed=30 + i, umax=3, vmax=2) for i in range(n)]
    rng = np.random.default_rng(4)
    mvs = [_prior(k, rng, H, W, pad=(0, 0)) for k in ("zero", "blocks", "int", "frac", "far")]
    I1 = _t(np.stack([f["I1"] for f in fps])); I2 = _t(np.stack([f["I2"] for f in fps])); mv = _t(np.stack(mvs))
    bD = torch.empty((n, H, W), dtype=torch.int32, device="cuda"); mC = torch.empty_like(bD)
    sub = torch.empty((n, 2, H, W), dtype=torch.float64, device="cuda")
    ctx.calc_pyd_cost_sgm_dev(I1, I2, mv, 5, 5, 2, 1, 6, 32, 1, 2, 0, bD, mC, sub)
    for i in range(n):
        want = _oracle_pyd(oracle, fps[i]["I1"], fps[i]["I2"], mvs[i], 5, 5, 2, 1, 6, 32, 1, 2, 0)
        assert np.array_equal(mC.cpu().numpy()[i].view(np.uint32), want["minC"]), i
        assert np.array_equal(bD.cpu().numpy()[i].view(np.uint32), want["bestD"]), i
        assert np.array_equal(sub.cpu().numpy()[i], want["mvSub"], equal_nan=True), i
