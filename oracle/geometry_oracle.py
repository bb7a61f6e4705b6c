"""CPU restatement of the dense epipolar prologue / epilogue of the MATLAB driver — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/rotation_motion.m:7-35 (+ computeEpipoleLineI2 :49-54), epipolar_geometry.m:104-119 and
epipolar_sgm_of.m:46-51.  MATLAB is not installed here and the reference holds no fixture for these arrays; MATLAB
evaluates F*P0 and H*P0 through BLAS (summation order and FMA use unspecified), so **bit parity with MATLAB is unpinned**.
The order is fixed as (m1*x + m2*y) + m3 with separately rounded numpy operations; every other step is an IEEE
elementwise operation (+, -, *, /, sqrt) that MATLAB and numpy round identically.
"""
from __future__ import annotations

import numpy as np


def _row(m, x, y):
    return (m[0] * x + m[1] * y) + m[2]


def epipolar_geometry(F, Hm, epi, direction, W, H):
    """-> Pd0 [2][H][W] (1-based), dirn [2][H][W], O [H][W], Rflow [2][H][W]"""
    F = np.asarray(F, np.float64).reshape(3, 3)
    Hm = np.asarray(Hm, np.float64).reshape(3, 3)
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)                       # rotation_motion.m:11-13 (0-based)
    with np.errstate(all="ignore"):
        l = [_row(F[i], x, y) for i in range(3)]                       # :16, :50
        nf = np.sqrt(l[0] * l[0] + l[1] * l[1])                        # :51
        nf = np.where(nf < 1e-6, 1.0, nf)                              # :52
        l = [v / nf for v in l]                                        # :53
        q = [_row(Hm[i], x, y) for i in range(3)]                      # :22
        p1 = [q[0] / q[2], q[1] / q[2], q[2] / q[2]]                   # :23
        coeff = -((l[0] * p1[0] + l[1] * p1[1]) + l[2] * p1[2])        # :28
        rx = (p1[0] - x) + coeff * l[0]                                # :24, :29
        ry = (p1[1] - y) + coeff * l[1]
        p0x, p0y = (x + 1.0) + rx, (y + 1.0) + ry                      # epipolar_geometry.m:107-114
        dx, dy = p0x - epi[0], p0y - epi[1]                            # :115
        if direction:
            dx, dy = -dx, -dy                                          # :116-118
        ln = np.sqrt(dx * dx + dy * dy)                                # :120
        dirn = np.stack([dx / ln, dy / ln])                            # :121
    return (np.ascontiguousarray(np.stack([p0x, p0y])), np.ascontiguousarray(dirn), np.ascontiguousarray(ln),
            np.ascontiguousarray(np.stack([rx, ry])))


def epipolar_flow(bestD, dirn, Rflow):
    """epipolar_sgm_of.m:46-49: flow = (bestD/256) .* normlizeDirection + flowR"""
    d = bestD.astype(np.float64) / 256.0
    return np.ascontiguousarray(d[None] * dirn + Rflow)
