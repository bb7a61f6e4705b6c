"""CPU suite: the C restatement (oracle/fsgm_oracle.c) against (i) the committed golden vectors produced by the
reference's own C++ and (ii) that C++ itself where it is present (authoring container), plus the parity-checklist
quirks of SURVEY.md §8a expressed as known-answer tests on the oracle."""
import glob
import os

import numpy as np
import pytest

from fsgm_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _g(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLD, "*.npz"))) >= 10


@pytest.mark.parametrize("name", ["epi_p8", "epi_p4", "epi_odd", "epi_wrap"])
def test_port_epi_vs_golden(oracle, name):
    g = _g(name)
    r = oracle.port_epi(g["I1"], g["I2"], int(g["D"]), float(g["vMax"]), g["Pd0"], g["dirn"], g["O"], int(g["P1"]), int(g["P2"]),
                        paths=int(g["paths"]))
    for k in ("cen1", "cen2", "Craw", "C", "minC", "bestD"):
        assert np.array_equal(r[k], g[k]), k
    assert np.array_equal(r["Sp"], g["Sp"].astype(np.uint32))


def test_port_fb_check_vs_golden(oracle):
    """forward_backward_check / calc_disp_from_first / convert_vzInd_to_disp (calc_cost_sgm.cpp:414-536)"""
    g = _g("epi_fb")
    D, vMax = int(g["D"]), float(g["vMax"])
    r = oracle.port_epi_fb(g["I1"], g["I2"], D, vMax, g["Pd0"], g["dirn"], g["O"], int(g["P1"]), int(g["P2"]), paths=8, thr=2)
    for k in ("bestD", "minC", "conf", "bestD2"):
        assert np.array_equal(r[k], g[k]), k
    conf, D2 = oracle.port_fb_check(g["D1"], g["Pd0"], g["dirn"], g["O"], vMax, D + 1, thr=int(g["thr_d"]))
    assert np.array_equal(conf, g["conf_d"]) and np.array_equal(D2, g["bestD2_d"])
    assert 0 < conf.mean() < 1 and (D2 == 512 << 8).any()                # the case exercises every branch
    assert np.array_equal(oracle.port_vz_to_disp(g["D1"], g["O"], vMax, D), g["disp_d"])


@pytest.mark.parametrize("name", ["pyd_a", "pyd_b", "pyd_c"])
def test_port_pyd_vs_golden(oracle, name):
    g = _g(name)
    r = oracle.port_pyd(g["I1"], g["I2"], g["preMv"], int(g["rx"]), int(g["ry"]), int(g["agg"]), int(g["sub"]), int(g["P1"]),
                        int(g["P2"]), int(g["diag"]), int(g["passes"]), int(g["adaptive"]))
    for k in ("C", "minC", "bestD", "mvSub"):
        assert np.array_equal(r[k], g[k]), k
    assert np.array_equal(r["Sp"], g["Sp"].astype(np.uint32))


def test_port_ng_vs_golden(oracle):
    g = _g("ng_a")
    r = oracle.port_ng(g["I1"], g["I2"], int(g["P1"]), int(g["P2"]), seed=int(g["seed"]), stages=True)
    assert np.array_equal(r["Centries"], g["Centries"].astype(np.int32))
    assert np.array_equal(r["Sp"], g["Sp"].astype(np.uint32))
    assert np.array_equal(r["minC"], g["minC"]) and np.array_equal(r["flow"], g["flow"])


@pytest.mark.parametrize("name", ["pydng_a", "pydng_b"])
def test_port_pydng_vs_golden(oracle, name):
    g = _g(name)
    r = oracle.port_pydng(g["I1"], g["I2"], g["preMv"], int(g["r"]), int(g["aggSize"]), int(g["sub"]), int(g["P1"]), int(g["P2"]),
                          stages=True)
    assert np.array_equal(r["Centries"], g["Centries"].astype(np.int32))
    assert np.array_equal(r["Sp"], g["Sp"].astype(np.uint32))
    assert np.array_equal(r["minC"], g["minC"]) and np.array_equal(r["flow"], g["flow"])


# ---- restatement vs the reference build itself (only where oracle/_ref exists) ------------------------------
def _need_ref(oracle, v):
    if not oracle.have_ref(v):
        pytest.skip("oracle/_ref not built here")


@pytest.mark.parametrize("W,H,D,thr,bad", [(64, 40, 16, 2, False), (50, 33, 64, 300, False), (40, 24, 8, 100, True)])
def test_port_vs_ref_fb_check(oracle, W, H, D, thr, bad):
    _need_ref(oracle, "epi8fb")
    p = synth.epipolar_pair(W, H, D, seed=W + 1)
    Pd0, dirn, O = p["Pd0"].copy(), p["dirn"].copy(), p["O"].copy()
    if bad:                                                  # NaN / huge geometry: x86 conversions give INT_MIN
        Pd0[0, 3, 5] = np.nan; dirn[1, 7, 7] = 1e300; O[9, 9] = -1e200; Pd0[1, 2, 2] = 3e9; O[11, 3] = np.inf
    D1 = np.random.default_rng(W).integers(0, D * 256, (H, W)).astype(np.uint32)
    a = oracle.ref_fb_check(D1, Pd0, dirn, O, p["vMax"], D + 1, thr)
    b = oracle.port_fb_check(D1, Pd0, dirn, O, p["vMax"], D + 1, thr)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(oracle.ref_vz_to_disp(D1, O, p["vMax"], D + 1), oracle.port_vz_to_disp(D1, O, p["vMax"], D))
    if not bad:
        r = oracle.ref_epi_fb(p["I1"], p["I2"], D, p["vMax"], Pd0, dirn, O, 6, 64)
        q = oracle.port_epi_fb(p["I1"], p["I2"], D, p["vMax"], Pd0, dirn, O, 6, 64, paths=8, thr=2)
        for k in ("bestD", "minC", "conf", "bestD2"):
            assert np.array_equal(r[k], q[k]), k


@pytest.mark.parametrize("W,H,D,P1,P2,paths", [(72, 40, 24, 6, 64, 8), (33, 21, 64, 6, 32, 4), (30, 20, 9, 250, 250, 8)])
def test_port_vs_ref_epi(oracle, W, H, D, P1, P2, paths):
    _need_ref(oracle, "epi8")
    p = synth.epipolar_pair(W, H, D, seed=W)
    a = oracle.ref_epi(p["I1"], p["I2"], D, p["vMax"], p["Pd0"], p["dirn"], p["O"], P1, P2, paths=paths)
    b = oracle.port_epi(p["I1"], p["I2"], D, p["vMax"], p["Pd0"], p["dirn"], p["O"], P1, P2, paths=paths)
    for k in ("cen1", "cen2", "Craw", "C", "Sp", "minC", "bestD"):
        assert np.array_equal(a[k], b[k]), k


def test_port_vs_ref_epi_wild_geometry(oracle):
    """NaN / huge / negative geometry exercises the x86 double->int behaviour."""
    _need_ref(oracle, "epi8")
    W, H, D = 24, 16, 8
    p = synth.epipolar_pair(W, H, D, seed=3)
    rng = np.random.default_rng(1)
    O = p["O"].copy()
    O[0, :4] = [np.nan, 1e300, -1e300, 3e9]
    O[1, :3] = [-5.0, 2.0 ** 31, -(2.0 ** 31) - 7]
    Pd0 = p["Pd0"] + rng.normal(0, 0.3, p["Pd0"].shape)
    Pd0[0, 2, :3] = [0.5, 1.5, -2.5]            # exact .5 ties after the -1
    a = oracle.ref_epi(p["I1"], p["I2"], D, 0.3, Pd0, p["dirn"], O, 6, 64, paths=8)
    b = oracle.port_epi(p["I1"], p["I2"], D, 0.3, Pd0, p["dirn"], O, 6, 64, paths=8)
    for k in ("Craw", "C", "Sp", "minC"):
        assert np.array_equal(a[k], b[k]), k


def test_port_vs_ref_pyd_ng(oracle):
    _need_ref(oracle, "pyd")
    fp = synth.flow_pair(40, 28, seed=2, umax=3, vmax=2)
    rng = np.random.default_rng(4)
    mv = rng.normal(0, 2.5, (2, 30, 44))
    a = oracle.ref_pyd(fp["I1"], fp["I2"], mv, 3, 2, 2, 1, 6, 32, 1, 2, 1)
    b = oracle.port_pyd(fp["I1"], fp["I2"], mv, 3, 2, 2, 1, 6, 32, 1, 2, 1)
    for k in ("C", "Sp", "minC", "bestD", "mvSub"):
        assert np.array_equal(a[k], b[k]), k
    a = oracle.ref_ng(fp["I1"], fp["I2"], 6, 32, seed=3, stages=True)
    b = oracle.port_ng(fp["I1"], fp["I2"], 6, 32, seed=3, stages=True)
    for k in ("Centries", "Sp", "minC", "flow"):
        assert np.array_equal(a[k], b[k]), k
    a = oracle.ref_pydng(fp["I1"], fp["I2"], mv, 1, 5, 1, 6, 32, stages=True)
    b = oracle.port_pydng(fp["I1"], fp["I2"], mv, 1, 5, 1, 6, 32, stages=True)
    for k in ("Centries", "Sp", "minC", "flow"):
        assert np.array_equal(a[k], b[k]), k


# ---- parity-checklist quirks as known answers ------------------------------------------------
def test_census_bit_layout(oracle):
    """tap k lands in bit 25-k; bit 0 always 0; centre tap always set; '>=' comparison; replicate border."""
    img = np.zeros((7, 7), np.uint8)
    img[3, 3] = 10
    cen = oracle.port_census(img)
    assert cen[3, 3] == 1 << 13                      # only the centre (k=12) is >= 10
    assert cen[0, 0] == ((1 << 25) - 1) << 1         # flat zero neighbourhood (3,3 is outside the window of 0,0)
    img[:] = 5
    assert np.all(oracle.port_census(img) == ((1 << 25) - 1) << 1)
    img[3, 1] = 9                                    # tap (dy=0,dx=-2) of pixel (3,3) -> k=10 -> bit 15
    assert (oracle.port_census(img)[3, 3] >> 15) & 1 == 1
    assert oracle.port_census(img)[3, 1] == 1 << 13


def test_wta_subpixel_quirks(oracle):
    """label 1 is never refined; label D-1 is refined against the NEXT pixel's label 0 (calc_cost_sgm.cpp:293-296)."""
    D = 8
    Sp = np.full((1, 3, D), 100, np.uint32)
    Sp[0, 0, 1] = 10; Sp[0, 0, 0] = 50; Sp[0, 0, 2] = 20          # argmin 1 -> 256 exactly, no refinement
    Sp[0, 1, 7] = 10; Sp[0, 1, 6] = 40                            # argmin 7 -> uses Sp[0,2,0]
    Sp[0, 2, 0] = 70
    Sp[0, 2, 3] = 10; Sp[0, 2, 2] = 30; Sp[0, 2, 4] = 20          # ordinary refinement
    bestD, minC = oracle.port_epi_wta(Sp)
    assert bestD[0, 0] == 256
    c_1, c, c1 = 40.0, 10.0, 70.0
    assert bestD[0, 1] == int((7 + (c1 - c_1) / (c - c1) / 2.0) * 256) != 7 * 256
    c_1, c, c1 = 30.0, 10.0, 20.0
    assert bestD[0, 2] == int((3 + (c1 - c_1) / (c - c_1) / 2.0) * 256)
    assert list(minC[0]) == [10, 10, 10]


def test_wta_first_minimum(oracle):
    Sp = np.full((1, 1, 6), 9, np.uint32)
    Sp[0, 0, 2] = Sp[0, 0, 4] = 3
    bestD, _ = oracle.port_epi_wta(Sp, subpixel=0)
    assert bestD[0, 0] == 2


def test_path_start_min_is_zero(oracle):
    """At a path start the stored minimum is 0, not min(C) (calc_cost_sgm.cpp:154): with C = 10 everywhere the
    second pixel gets L = 10 + min(10, 0 + P2) - 0 = 20, not 10."""
    Cv = np.full((1, 3, 4), 10, np.uint8)
    I1 = np.zeros((1, 3), np.uint8)
    L = oracle.port_sweep1d(Cv, I1, 6, 64, 0)
    assert L[0, 0].tolist() == [10] * 4 and L[0, 1].tolist() == [20] * 4 and L[0, 2].tolist() == [10] * 4


def test_u8_wraparound(oracle):
    """PathCost is unsigned char with wrap-around (common.h:4-8): LpreMin + P2 wraps when it exceeds 255."""
    Cv = np.zeros((1, 3, 3), np.uint8)
    Cv[0, 0] = [200, 250, 250]
    Cv[0, 1] = [5, 5, 5]
    I1 = np.zeros((1, 3), np.uint8)
    L = oracle.port_sweep1d(Cv, I1, 10, 100, 0)
    # pixel 1: preMin = 0 (path start) so far = 100: L = 5 + min(100, Lpre[d], nb+10 mod 256) - 0
    assert L[0, 1].tolist() == [5 + min(100, 200, (250 + 10) % 256), 5 + min(100, 250, 210, 4), 5 + min(100, 250, 4)]


def test_labels_use_D_plus_1(oracle):
    assert synth.vz_index(5, 9, 0.3) == (1.0 * 5 / 10 * 0.3) / (1 - 1.0 * 5 / 10 * 0.3)
