"""Seeded synthetic inputs for the fSGM hot path (SURVEY.md §8d recipes A-E).

numpy only; used by tests/, bench.py and __graft_entry__.smoke().  Every array is
row-major with x fastest, exactly what the reference's gateways expect after the
MATLAB side has permuted its arrays (epipolar_sgm_of.m:33-43): images are uint8
[H][W]; two-plane double arrays are plane-major [2][H][W] (plane 0 = X).
"""
from __future__ import annotations

import numpy as np


def texture(W: int, H: int, seed: int) -> np.ndarray:
    """Band-limited noise image (uint8 [H][W]) so that the 5x5 census is informative."""
    rng = np.random.default_rng(seed)
    img = np.zeros((H, W), np.float64)
    for octave, amp in ((1, 0.45), (2, 0.3), (4, 0.25)):
        h, w = (H + octave - 1) // octave + 2, (W + octave - 1) // octave + 2
        n = rng.random((h, w))
        up = np.kron(n, np.ones((octave, octave)))[:H, :W]
        if octave > 1:  # cheap blur so coarse octaves are smooth
            k = octave
            c = np.cumsum(np.pad(up, ((k, k), (k, k)), mode="edge"), 0)
            up = (c[2 * k:, :] - c[:-2 * k, :])[:, k:-k] / (2 * k)
            c = np.cumsum(np.pad(up, ((0, 0), (k, k)), mode="edge"), 1)
            up = (c[:, 2 * k:] - c[:, :-2 * k]) / (2 * k)
        img += amp * up[:H, :W]
    img = (img - img.min()) / (img.max() - img.min() + 1e-12)
    return np.ascontiguousarray((img * 255.0 + 0.5).astype(np.uint8))


def smooth_field(W: int, H: int, lo: float, hi: float, seed: int, cells: int = 6) -> np.ndarray:
    """Smooth scalar field in [lo, hi] (bilinear interpolation of a coarse random grid)."""
    rng = np.random.default_rng(seed)
    g = rng.random((cells + 1, cells + 1))
    ys = np.linspace(0, cells, H)
    xs = np.linspace(0, cells, W)
    y0 = np.clip(ys.astype(int), 0, cells - 1)
    x0 = np.clip(xs.astype(int), 0, cells - 1)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    f = (g[y0][:, x0] * (1 - fy) * (1 - fx) + g[y0][:, x0 + 1] * (1 - fy) * fx
         + g[y0 + 1][:, x0] * fy * (1 - fx) + g[y0 + 1][:, x0 + 1] * fy * fx)
    return lo + (hi - lo) * f


def vz_index(d, D: int, vMax: float):
    """Label -> vz index, the reference's exact expression (calc_cost_sgm.cpp:339,360-361)."""
    n = float(D + 1)
    r = 1.0 * np.asarray(d, np.float64) / n * vMax
    return r / (1 - r)


def epipolar_pair(W: int, H: int, D: int, seed: int = 1, vMax: float = 0.3):
    """Recipes A/B: texture I1, pure forward translation (epipole near the centre, no rotation),
    smooth ground-truth label field, I2 forward-splatted from I1.

    Returns dict(I1, I2, Pd0 [2][H][W] 1-based, dirn [2][H][W], O [H][W], vMax, D, gt_label).
    """
    I1 = texture(W, H, seed)
    ex, ey = W / 2 + 0.37, H / 2 + 0.21
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    vx, vy = xx - ex, yy - ey
    O = np.sqrt(vx * vx + vy * vy)
    dirn = np.stack([vx / O, vy / O])
    Pd0 = np.stack([xx + 1.0, yy + 1.0])
    gt = np.floor(smooth_field(W, H, 0.0, D - 1e-6, seed + 1)).astype(np.int64)
    off = O * vz_index(gt, D, vMax)
    x2 = np.clip(np.round(xx + off * dirn[0]).astype(np.int64), 0, W - 1)
    y2 = np.clip(np.round(yy + off * dirn[1]).astype(np.int64), 0, H - 1)
    I2 = texture(W, H, seed + 7919)          # background for pixels nothing maps to
    I2[y2, x2] = I1                            # forward splat (last writer wins; fine for a test pattern)
    return dict(I1=I1, I2=np.ascontiguousarray(I2), Pd0=np.ascontiguousarray(Pd0),
                dirn=np.ascontiguousarray(dirn), O=np.ascontiguousarray(O), vMax=vMax, D=D, gt_label=gt)


def flow_pair(W: int, H: int, seed: int = 1, umax: int = 20, vmax: int = 10, blocks: int = 4):
    """Recipe C/D: piecewise-constant integer true flow; I2[y+v, x+u] = I1[y, x]."""
    I1 = texture(W, H, seed)
    rng = np.random.default_rng(seed + 3)
    bu = rng.integers(-umax, umax + 1, (blocks, blocks))
    bv = rng.integers(-vmax, vmax + 1, (blocks, blocks))
    by = np.minimum((np.arange(H) * blocks) // H, blocks - 1)
    bx = np.minimum((np.arange(W) * blocks) // W, blocks - 1)
    u = bu[by][:, bx]
    v = bv[by][:, bx]
    yy, xx = np.mgrid[0:H, 0:W]
    I2 = texture(W, H, seed + 7919)
    I2[np.clip(yy + v, 0, H - 1), np.clip(xx + u, 0, W - 1)] = I1
    return dict(I1=I1, I2=np.ascontiguousarray(I2), u=u.astype(np.float64), v=v.astype(np.float64))


def reduce2(img: np.ndarray) -> np.ndarray:
    """Harness pyramid step (SURVEY.md §8d recipe C): 5-tap [1 4 6 4 1]/16 blur then 2x decimation,
    output size ceil(/2) like impyramid 'reduce' (pyramidal_sgm.m:29-30).  Not a parity target:
    the same arrays are fed to oracle and GPU."""
    k = np.array([1, 4, 6, 4, 1], np.float64) / 16
    p = np.pad(img.astype(np.float64), 2, mode="edge")
    t = sum(k[i] * p[:, i:i + img.shape[1]] for i in range(5))
    t = sum(k[i] * t[i:i + img.shape[0], :] for i in range(5))
    return np.ascontiguousarray(np.clip(t[::2, ::2] + 0.5, 0, 255).astype(np.uint8))


def upsample_mv(mv: np.ndarray, Hn: int, Wn: int) -> np.ndarray:
    """2*nearest-neighbour x2 upsample of a [2][h][w] mv map (pyramidal_sgm.m:72), cropped/padded to >= (Hn, Wn)."""
    up = 2.0 * np.repeat(np.repeat(mv, 2, axis=1), 2, axis=2)
    out = np.zeros((2, max(Hn, up.shape[1]), max(Wn, up.shape[2])), np.float64)
    out[:, :up.shape[1], :up.shape[2]] = up
    return np.ascontiguousarray(out)


def epipolar_camera(W: int, H: int, seed: int = 1, rot_deg: float = 0.4):
    """Two-view geometry for the dense prologue: pinhole K, a small rotation R and a mostly-forward
    translation t.  Returns dict(F [9], H [9] = K R K^-1 (row-major), epi [2] (1-based pixel position of the epipole in
    image 2, as MATLAB's epi(1:2)), direction)."""
    rng = np.random.default_rng(seed)
    f = 0.9 * W
    K = np.array([[f, 0, W / 2 + 0.37], [0, f, H / 2 + 0.21], [0, 0, 1.0]])
    a = np.deg2rad(rot_deg) * rng.uniform(-1, 1, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    R = Rz @ Ry @ Rx
    t = np.array([0.03 * rng.uniform(-1, 1), 0.02 * rng.uniform(-1, 1), 1.0])
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Ki = np.linalg.inv(K)
    F = Ki.T @ tx @ R @ Ki
    F = F / np.abs(F).max()
    Hm = K @ R @ Ki
    e = K @ t
    epi = np.array([e[0] / e[2] + 1.0, e[1] / e[2] + 1.0])
    return dict(F=np.ascontiguousarray(F.reshape(9)), H=np.ascontiguousarray(Hm.reshape(9)), epi=epi, direction=int(seed % 2))
