"""CPU restatement of the MATLAB pyramid driver around calc_pyd_cost_sgm — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/pyramidal_sgm.m:24-75.  The per-level solver is the pinned oracle (oracle/_ref build of
calc_pyd_cost_sgm.cpp, or the C restatement).  ``impyramid`` is MATLAB Image Processing Toolbox code and is NOT in the
reference tree, MATLAB/Octave are not installed here, and the reference holds no image-pyramid fixture:
**parity of the 'reduce' step is unpinned**.  It is restated below from the toolbox's published algorithm
(impyramid.m -> imresize.m with a custom kernel):

  * kernel = piecewise-constant function with values [.0625 .25 .375 .25 .0625] on the half-open unit intervals
    ending at 2.5, 1.5, 0.5, -0.5, -1.5 (Burt-Adelson a = 0.375), kernel width 5, scale 1/2, output size ceil(n/2),
    'Antialiasing' false;
  * imresize's contributions(): for 1-based output index x, u = x/scale + 0.5*(1 - 1/scale); left = floor(u - width/2);
    indices = left + (0 .. ceil(width)+1); weights = kernel(u - indices), normalised to sum 1; out-of-range indices are
    looked up in aux = [1..n, n..1] (symmetric border);
  * dimensions are resized in order of increasing scale (equal scales: dim 1 = MATLAB rows = image y first), and for
    uint8 input the MEX kernel saturates and rounds (half away from zero) to uint8 after EACH pass.

`reduce_general` evaluates that recipe literally in floating point; `reduce_closed_form` is the integer closed form the
CUDA kernel uses ((sum+8)>>4 per pass, taps 2o-2..2o+2, mirror border) — tests check the two agree.
"""
from __future__ import annotations

import numpy as np


def _kernel(x):
    brk = [3.5, 2.5, 1.5, 0.5, -0.5, -1.5, -np.inf]
    val = [0.0, 0.0625, 0.25, 0.375, 0.25, 0.0625, 0.0]
    out = np.zeros_like(x, dtype=np.float64)
    for i, xv in np.ndenumerate(x):
        for b, v in zip(brk, val):
            if xv >= b:
                out[i] = v
                break
    return out


def _contributions(n_in, n_out, scale=0.5, width=5.0):
    x = np.arange(1, n_out + 1, dtype=np.float64)
    u = x / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - width / 2)
    P = int(np.ceil(width)) + 2
    ind = left[:, None] + np.arange(P)[None, :]
    w = _kernel(u[:, None] - ind)
    w = w / w.sum(axis=1, keepdims=True)
    aux = np.concatenate([np.arange(n_in), np.arange(n_in - 1, -1, -1)])
    ind = aux[np.mod(ind.astype(np.int64) - 1, aux.size)]
    return w, ind


def _round_u8(v):
    return np.clip(np.floor(v + 0.5), 0, 255).astype(np.uint8)


def reduce_general(img: np.ndarray) -> np.ndarray:
    H, W = img.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    wy, iy = _contributions(H, Ho)
    tmp = _round_u8(np.einsum("ok,okx->ox", wy, img[iy, :].astype(np.float64)))           # dim 1 (y) first
    wx, ix = _contributions(W, Wo)
    return np.ascontiguousarray(_round_u8(np.einsum("ok,yok->yo", wx, tmp[:, ix].astype(np.float64))))


def _mirror(i, n):
    m = np.mod(i, 2 * n)
    return np.where(m < n, m, 2 * n - 1 - m)


def reduce_closed_form(img: np.ndarray) -> np.ndarray:
    H, W = img.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    wgt = np.array([1, 4, 6, 4, 1], np.uint32)
    iy = _mirror(2 * np.arange(Ho)[:, None] - 2 + np.arange(5)[None, :], H)
    tmp = ((img[iy, :].astype(np.uint32) * wgt[None, :, None]).sum(1) + 8) >> 4
    ix = _mirror(2 * np.arange(Wo)[:, None] - 2 + np.arange(5)[None, :], W)
    return np.ascontiguousarray((((tmp[:, ix] * wgt[None, None, :]).sum(2) + 8) >> 4).astype(np.uint8))


def pyramid(img: np.ndarray, levels: int):
    out = [np.ascontiguousarray(img)]
    for _ in range(1, levels):
        out.append(reduce_closed_form(out[-1]))
    return out


def pyramidal_sgm(I0, I1, solver, numPyd=5, P1=6, P2=32, agg=2, ver=5, hor=5, diag=1, passes=2, adaptive=0):
    """pyramidal_sgm.m:24-75.  `solver(I1, I2, preMv, rx, ry, agg, sub, P1, P2, diag, passes, adaptive)` -> dict with
    bestD, minC, mvSub (oracle.ref_pyd / oracle.port_pyd).  Returns (mv [2][H][W], minC, [mv per level, finest first])."""
    p0, p1 = pyramid(I0, numPyd), pyramid(I1, numPyd)
    Hc, Wc = p0[-1].shape
    pre = np.zeros((2, Hc, Wc))                                       # :33
    per_level = [None] * numPyd
    minC = None
    for l in range(numPyd - 1, -1, -1):                               # :36
        H, W = p0[l].shape
        r = solver(p0[l], p1[l], np.ascontiguousarray(pre), hor, ver, agg, int(l == 0), P1, P2, diag, passes, adaptive)
        idx = r["bestD"].astype(np.int64)                             # label = sx*Sy + sy; ind2sub over [Sy, Sx] (:57)
        sx, sy = idx // (2 * ver + 1), idx % (2 * ver + 1)
        lab = np.stack([sx - hor, sy - ver]).astype(np.float64)       # :59-60
        mv = lab + pre[:, :H, :W] + r["mvSub"]                        # :64
        per_level[l] = mv
        minC = r["minC"]
        if l > 0:
            pre = 2.0 * np.repeat(np.repeat(mv, 2, axis=1), 2, axis=2)    # :72  2*imresize(mv, 2, 'nearest')
    return per_level[0], minC, per_level
