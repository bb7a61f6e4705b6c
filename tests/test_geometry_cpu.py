"""CPU suite for the dense epipolar prologue / epilogue restatement (rotation_motion.m, epipolar_geometry.m:104-119,
epipolar_sgm_of.m:46-51).  MATLAB is absent, so these are known-answer and invariance checks."""
import numpy as np

from fsgm_b200 import synth


def test_pure_translation_reduces_to_recipe_a():
    """H = I and F = [e]_x (pure translation): every pixel lies on its own epipolar line, so Rflow vanishes and the maps are
    the ones synth.epipolar_pair writes down directly (Pd0 = p + 1, direction = (p - e)/|p - e|, offset = |p - e|)."""
    from oracle import geometry_oracle as go
    W, H = 64, 40
    e = np.array([W / 2 + 0.37, H / 2 + 0.21, 1.0])                     # 0-based epipole
    F = np.array([[0, -e[2], e[1]], [e[2], 0, -e[0]], [-e[1], e[0], 0]])
    Pd0, dirn, O, R = go.epipolar_geometry(F, np.eye(3), e[:2] + 1.0, 0, W, H)
    p = synth.epipolar_pair(W, H, 16, seed=1)
    assert np.abs(R).max() < 1e-9
    assert np.allclose(Pd0, p["Pd0"], atol=1e-9) and np.allclose(dirn, p["dirn"], atol=1e-9) and np.allclose(O, p["O"], atol=1e-9)
    # direction flag negates the direction, not the offset
    _, dneg, Oneg, _ = go.epipolar_geometry(F, np.eye(3), e[:2] + 1.0, 1, W, H)
    assert np.array_equal(dneg, -dirn) and np.array_equal(Oneg, O)


def test_zero_disparity_point_lies_on_the_epipolar_line():
    from oracle import geometry_oracle as go
    W, H = 90, 60
    c = synth.epipolar_camera(W, H, seed=4, rot_deg=1.0)
    Pd0, dirn, O, R = go.epipolar_geometry(c["F"], c["H"], c["epi"], c["direction"], W, H)
    F = c["F"].reshape(3, 3)
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    l = np.stack([F[i, 0] * x + F[i, 1] * y + F[i, 2] for i in range(3)])
    l /= np.sqrt(l[0] ** 2 + l[1] ** 2)
    resid = l[0] * (Pd0[0] - 1) + l[1] * (Pd0[1] - 1) + l[2]            # PrefD0 (0-based) on l = F p
    assert np.abs(resid).max() < 1e-8
    assert np.allclose(dirn[0] ** 2 + dirn[1] ** 2, 1.0) and np.abs(R).max() > 0.05
    flow = go.epipolar_flow(np.full((H, W), 512, np.uint32), dirn, R)   # disparity 2 px everywhere
    assert np.allclose(flow, 2.0 * dirn + R)


def test_degenerate_line_normalisation():
    """|l(1:2)| < 1e-6 keeps the line unnormalised (rotation_motion.m:52)"""
    from oracle import geometry_oracle as go
    F = np.zeros((3, 3)); F[2, 2] = 1.0
    Pd0, dirn, O, R = go.epipolar_geometry(F, np.eye(3), [5.0, 4.0], 0, 8, 6)
    y, x = np.mgrid[0:6, 0:8].astype(np.float64)
    assert np.array_equal(R, np.zeros_like(R))                          # coeff = -1 times l(1:2) = 0
    assert np.array_equal(Pd0, np.stack([x + 1, y + 1]))
