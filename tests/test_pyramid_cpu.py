"""CPU suite for the pyramid driver (pyramidal_sgm.m): the numpy restatement against its own literal
imresize recipe, against the committed fixture (per-level solver = the reference's C++) and the C port."""
import os

import numpy as np
import pytest

from fsgm_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _g(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("H,W", [(37, 50), (1, 7), (7, 1), (2, 2), (3, 4), (94, 311), (5, 5), (64, 64)])
def test_reduce_closed_form_matches_imresize_recipe(H, W):
    from oracle import pyramid_oracle as pyo
    img = np.random.default_rng(H * 1000 + W).integers(0, 256, (H, W)).astype(np.uint8)
    a, b = pyo.reduce_general(img), pyo.reduce_closed_form(img)
    assert a.shape == ((H + 1) // 2, (W + 1) // 2) and np.array_equal(a, b)


def test_reduce_known_answers():
    """constant images stay constant; an impulse spreads as [1 4 6 4 1]/16 with per-pass uint8 rounding; the border repeats the edge"""
    from oracle import pyramid_oracle as pyo
    assert np.array_equal(pyo.reduce_closed_form(np.full((9, 12), 201, np.uint8)), np.full((5, 6), 201, np.uint8))
    img = np.zeros((9, 9), np.uint8)
    img[4, 4] = 255
    out = pyo.reduce_closed_form(img)
    col = [(255 * w + 8) >> 4 for w in (1, 6, 1)]                 # vertical pass at rows 2o = 2, 4, 6 -> 16, 96, 16
    assert [int(v) for v in out[1:4, 2]] == [(c * 6 + 8) >> 4 for c in col]
    assert int(out[2, 1]) == (col[1] * 1 + 8) >> 4 and int(out[2, 3]) == (col[1] * 1 + 8) >> 4
    edge = np.zeros((4, 6), np.uint8)
    edge[0, 0] = 160                                              # taps -2..2 at the corner: indices 1,0,0,1,2 -> weight 4+6 = 10
    assert int(pyo.reduce_closed_form(edge)[0, 0]) == ((((160 * 10 + 8) >> 4) * 10 + 8) >> 4)


def test_pyramid_driver_vs_golden(oracle):
    from oracle import pyramid_oracle as pyo
    g = _g("pyramid_a")
    L, ver, hor = int(g["numPyd"]), int(g["ver"]), int(g["hor"])
    for i, im in enumerate(pyo.pyramid(g["I0"], L)):
        assert np.array_equal(im, g[f"img_l{i}"])
    mv, mc, lv = pyo.pyramidal_sgm(g["I0"], g["I1"], lambda *a: oracle.port_pyd(*a, stages=False), numPyd=L, ver=ver, hor=hor)
    assert np.array_equal(mv, g["mv"]) and np.array_equal(mc, g["minC"])
    for i in range(L):
        assert np.array_equal(lv[i], g[f"mv_l{i}"]), i
    # a global shift is recovered on most pixels (sanity of the label -> mv, prior and upsampling conventions)
    fp = synth.flow_pair(90, 58, seed=21, umax=9, vmax=5, blocks=1)
    mv, _, _ = pyo.pyramidal_sgm(fp["I1"], fp["I2"], lambda *a: oracle.port_pyd(*a, stages=False), numPyd=L, ver=ver, hor=hor)
    err = np.maximum(np.abs(mv[0] - fp["u"]), np.abs(mv[1] - fp["v"]))
    assert (err <= 1.0).mean() > 0.85


def test_pyramid_dims_abi():
    from fsgm_b200 import api
    ws, hs = api.pyramid_dims(1242, 375, 5)
    assert ws == [1242, 621, 311, 156, 78] and hs == [375, 188, 94, 47, 24]
    with pytest.raises(api.FsgmError):
        api.pyramid_dims(10, 10, 17)
