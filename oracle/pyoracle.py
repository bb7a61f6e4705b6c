"""ctypes access to the CPU checkers — TEST INFRASTRUCTURE ONLY.

Two back ends with the same Python signatures:
  * ``ref``  : oracle/_ref/libref_*.so — the reference's own C++ compiled from /root/reference
               (built only in the authoring container; the .so files travel to the GPU box).
  * ``port`` : oracle/libfsgm_oracle.so — the plain-C restatement (fsgm_oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under fsgm_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u8p, _u32p, _i32p, _f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_int32, C.c_double))


def _ptr(a, ty):
    if a is None:
        return ty()
    if not a.flags["C_CONTIGUOUS"]:
        raise TypeError("oracle arrays must be C-contiguous")
    return a.ctypes.data_as(ty)


def build(port: bool = True, ref: bool = True) -> None:
    """(Re)build the checkers.  `ref` is a no-op where /root/reference is absent."""
    targets = (["port"] if port else []) + (["ref"] if ref else [])
    if targets:
        subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def have_ref(variant: str = "epi8") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libref_{variant}.so"))


_libs: dict = {}


def _ref(variant: str):
    if variant not in _libs:
        _libs[variant] = C.CDLL(os.path.join(HERE, "_ref", f"libref_{variant}.so"))
    return _libs[variant]


def _port():
    if "port" not in _libs:
        path = os.path.join(HERE, "libfsgm_oracle.so")
        if not os.path.exists(path):
            build(port=True, ref=False)
        _libs["port"] = C.CDLL(path)
    return _libs["port"]


# --------------------------------------------------------------------------------------
# reference back end
# --------------------------------------------------------------------------------------
def ref_epi(I1, I2, D, vMax, Pd0, dirn, O, P1, P2, paths=8, stages=True):
    """Runs calc_cost_sgm.cpp's mexFunction (4-path as shipped, or the 8-path build)."""
    H, W = I1.shape
    N = W * H
    lib = _ref("epi8" if paths == 8 else "epi4")
    lib.ref_epi.restype = C.c_double
    out = dict(bestD=np.empty((H, W), np.uint32), minC=np.empty((H, W), np.uint32))
    if stages:
        out.update(cen1=np.empty((H, W), np.uint32), cen2=np.empty((H, W), np.uint32),
                   Craw=np.empty((H, W, D), np.uint8), C=np.empty((H, W, D), np.uint8),
                   Sp=np.empty((H, W, D), np.uint32))
    g = out.get
    out["seconds"] = lib.ref_epi(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, D, C.c_double(vMax),
                                 _ptr(Pd0, _f64p), _ptr(dirn, _f64p), _ptr(O, _f64p), P1, P2,
                                 _ptr(out["bestD"], _u32p), _ptr(out["minC"], _u32p),
                                 _ptr(g("cen1"), _u32p), _ptr(g("cen2"), _u32p), _ptr(g("Craw"), _u8p),
                                 _ptr(g("C"), _u8p), _ptr(g("Sp"), _u32p))
    return out


def ref_epi_stage_times(I1, I2, D, vMax, Pd0, dirn, O, P1, P2, paths=8):
    H, W = I1.shape
    lib = _ref("epi8" if paths == 8 else "epi4")
    lib.ref_epi_stage_times.restype = C.c_double
    tc, ts = C.c_double(), C.c_double()
    tot = lib.ref_epi_stage_times(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, D, C.c_double(vMax), _ptr(Pd0, _f64p),
                                  _ptr(dirn, _f64p), _ptr(O, _f64p), P1, P2, C.byref(tc), C.byref(ts))
    return dict(total=tot, cost=tc.value, sgm=ts.value)


def ref_epi_fb(I1, I2, D, vMax, Pd0, dirn, O, P1, P2):
    """8-path build with the commented-out forward_backward_check call (calc_cost_sgm.cpp:589-590) re-enabled."""
    H, W = I1.shape
    lib = _ref("epi8fb")
    lib.ref_epi_all.restype = C.c_double
    out = dict(bestD=np.empty((H, W), np.uint32), minC=np.empty((H, W), np.uint32),
               conf=np.empty((H, W), np.uint8), bestD2=np.empty((H, W), np.uint32))
    out["seconds"] = lib.ref_epi_all(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, D, C.c_double(vMax), _ptr(Pd0, _f64p),
                                     _ptr(dirn, _f64p), _ptr(O, _f64p), P1, P2, _ptr(out["bestD"], _u32p),
                                     _ptr(out["minC"], _u32p), _ptr(out["conf"], _u8p), _ptr(out["bestD2"], _u32p))
    return out


def ref_fb_check(D1, Pd0, dirn, O, vMax, n, thr=2):
    """forward_backward_check (calc_cost_sgm.cpp:488-536) called directly on a x256 label map."""
    H, W = D1.shape
    conf, D2 = np.empty((H, W), np.uint8), np.empty((H, W), np.uint32)
    _ref("epi4").ref_fb_check(_ptr(D1, _u32p), W, H, _ptr(Pd0, _f64p), _ptr(dirn, _f64p), _ptr(O, _f64p),
                              C.c_double(vMax), n, thr, _ptr(conf, _u8p), _ptr(D2, _u32p))
    return conf, D2


def ref_vz_to_disp(D1, O, vMax, n):
    H, W = D1.shape
    out = np.ascontiguousarray(D1).copy()
    _ref("epi4").ref_vz_to_disp(_ptr(out, _u32p), W, H, _ptr(O, _f64p), C.c_double(vMax), n)
    return out


def ref_census(I):
    H, W = I.shape
    cen = np.empty((H, W), np.uint32)
    _ref("epi4").ref_census(_ptr(I, _u8p), _ptr(cen, _u32p), W, H)
    return cen


def ref_pyd(I1, I2, preMv, rx, ry, agg, subpix, P1, P2, diag=1, passes=2, adaptive=0, stages=True):
    H, W = I1.shape
    _, mvH, mvW = preMv.shape
    D = (2 * rx + 1) * (2 * ry + 1)
    lib = _ref("pyd")
    lib.ref_pyd.restype = C.c_double
    out = dict(bestD=np.empty((H, W), np.uint32), minC=np.empty((H, W), np.uint32), mvSub=np.empty((2, H, W), np.float64))
    if stages:
        out.update(cen1=np.empty((H, W), np.uint32), cen2=np.empty((H, W), np.uint32),
                   C=np.empty((H, W, D), np.uint8), Sp=np.empty((H, W, D), np.uint32))
    g = out.get
    out["seconds"] = lib.ref_pyd(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, _ptr(preMv, _f64p), mvW, mvH, rx, ry, agg, subpix,
                                 P1, P2, diag, passes, adaptive, _ptr(out["bestD"], _u32p), _ptr(out["minC"], _u32p),
                                 _ptr(out["mvSub"], _f64p), _ptr(g("cen1"), _u32p), _ptr(g("cen2"), _u32p),
                                 _ptr(g("C"), _u8p), _ptr(g("Sp"), _u32p))
    return out


def ref_ng(I1, I2, P1, P2, seed=1, stages=False):
    H, W = I1.shape
    D = 108
    lib = _ref("ng")
    lib.ref_ng.restype = C.c_double
    out = dict(minC=np.empty((H, W), np.uint32), flow=np.empty((2, H, W), np.float64))
    if stages:
        out.update(Centries=np.empty((H, W, D, 3), np.int32), Sp=np.empty((H, W, D), np.uint32))
    g = out.get
    out["seconds"] = lib.ref_ng(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, P1, P2, C.c_uint(seed), _ptr(out["minC"], _u32p),
                                _ptr(out["flow"], _f64p), _ptr(g("Centries"), _i32p), _ptr(g("Sp"), _u32p))
    return out


def ref_pydng(I1, I2, preMv, r, aggSize, subpix, P1, P2, stages=False):
    H, W = I1.shape
    _, mvH, mvW = preMv.shape
    D = 9 * (2 * r + 1) ** 2
    lib = _ref("pydng")
    lib.ref_pydng.restype = C.c_double
    out = dict(minC=np.empty((H, W), np.uint32), flow=np.empty((2, H, W), np.float64))
    if stages:
        out.update(Centries=np.empty((H, W, D, 3), np.int32), Sp=np.empty((H, W, D), np.uint32))
    g = out.get
    out["seconds"] = lib.ref_pydng(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, _ptr(preMv, _f64p), mvW, mvH, r, aggSize, subpix,
                                   P1, P2, _ptr(out["minC"], _u32p), _ptr(out["flow"], _f64p),
                                   _ptr(g("Centries"), _i32p), _ptr(g("Sp"), _u32p))
    return out


# --------------------------------------------------------------------------------------
# port back end (oracle/fsgm_oracle.c) — same return dictionaries as the ref_* functions
# --------------------------------------------------------------------------------------
def port_census(I):
    H, W = I.shape
    cen = np.empty((H, W), np.uint32)
    _port().orc_census(_ptr(I, _u8p), _ptr(cen, _u32p), W, H)
    return cen


def port_epi(I1, I2, D, vMax, Pd0, dirn, O, P1, P2, paths=8, stages=True):
    H, W = I1.shape
    out = dict(bestD=np.empty((H, W), np.uint32), minC=np.empty((H, W), np.uint32))
    if stages:
        out.update(cen1=np.empty((H, W), np.uint32), cen2=np.empty((H, W), np.uint32),
                   Craw=np.empty((H, W, D), np.uint8), C=np.empty((H, W, D), np.uint8),
                   Sp=np.empty((H, W, D), np.uint32))
    g = out.get
    _port().orc_epi(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, D, C.c_double(vMax), _ptr(Pd0, _f64p), _ptr(dirn, _f64p),
                    _ptr(O, _f64p), P1, P2, paths, _ptr(out["bestD"], _u32p), _ptr(out["minC"], _u32p),
                    _ptr(g("cen1"), _u32p), _ptr(g("cen2"), _u32p), _ptr(g("Craw"), _u8p), _ptr(g("C"), _u8p),
                    _ptr(g("Sp"), _u32p))
    return out


def port_fb_check(D1, Pd0, dirn, O, vMax, n, thr=2, use_vzind=1):
    H, W = D1.shape
    conf, D2 = np.empty((H, W), np.uint8), np.empty((H, W), np.uint32)
    _port().orc_fb_check(_ptr(D1, _u32p), W, H, _ptr(Pd0, _f64p), _ptr(dirn, _f64p), _ptr(O, _f64p),
                         C.c_double(vMax), n, thr, use_vzind, _ptr(conf, _u8p), _ptr(D2, _u32p))
    return conf, D2


def port_vz_to_disp(D1, O, vMax, D):
    H, W = D1.shape
    out = np.ascontiguousarray(D1).copy()
    _port().orc_vz_to_disp(_ptr(out, _u32p), W, H, _ptr(O, _f64p), C.c_double(vMax), D)
    return out


def port_epi_fb(I1, I2, D, vMax, Pd0, dirn, O, P1, P2, paths=8, thr=2):
    H, W = I1.shape
    out = dict(bestD=np.empty((H, W), np.uint32), minC=np.empty((H, W), np.uint32),
               conf=np.empty((H, W), np.uint8), bestD2=np.empty((H, W), np.uint32))
    _port().orc_epi_fb(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, D, C.c_double(vMax), _ptr(Pd0, _f64p), _ptr(dirn, _f64p),
                       _ptr(O, _f64p), P1, P2, paths, thr, _ptr(out["bestD"], _u32p), _ptr(out["minC"], _u32p),
                       _ptr(out["conf"], _u8p), _ptr(out["bestD2"], _u32p))
    return out


def port_sweep1d(Cvol, I1, P1, P2, r, adaptive_thr=0):
    """One direction r (0..7, order L1,L3,L2,L4 then reversed) -> L volume uint8 [H][W][D]."""
    H, W, D = Cvol.shape
    L = np.empty((H, W, D), np.uint8)
    _port().orc_sweep1d(_ptr(Cvol, _u8p), _ptr(I1, _u8p), W, H, D, P1, P2, adaptive_thr, r, _ptr(L, _u8p))
    return L


def port_epi_wta(Sp, subpixel=1):
    H, W, D = Sp.shape
    bestD, minC = np.empty((H, W), np.uint32), np.empty((H, W), np.uint32)
    _port().orc_epi_wta(_ptr(Sp, _u32p), W, H, D, subpixel, _ptr(bestD, _u32p), _ptr(minC, _u32p))
    return bestD, minC


def port_pyd(I1, I2, preMv, rx, ry, agg, subpix, P1, P2, diag=1, passes=2, adaptive=0, stages=True):
    H, W = I1.shape
    _, mvH, mvW = preMv.shape
    D = (2 * rx + 1) * (2 * ry + 1)
    out = dict(bestD=np.empty((H, W), np.uint32), minC=np.empty((H, W), np.uint32), mvSub=np.empty((2, H, W), np.float64))
    if stages:
        out.update(cen1=np.empty((H, W), np.uint32), cen2=np.empty((H, W), np.uint32),
                   C=np.empty((H, W, D), np.uint8), Sp=np.empty((H, W, D), np.uint32))
    g = out.get
    _port().orc_pyd(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, _ptr(preMv, _f64p), mvW, mvH, rx, ry, agg, subpix, P1, P2,
                    diag, passes, adaptive, _ptr(out["bestD"], _u32p), _ptr(out["minC"], _u32p), _ptr(out["mvSub"], _f64p),
                    _ptr(g("cen1"), _u32p), _ptr(g("cen2"), _u32p), _ptr(g("C"), _u8p), _ptr(g("Sp"), _u32p))
    return out


def port_ng(I1, I2, P1, P2, seed=1, stages=False):
    H, W = I1.shape
    D = 108
    out = dict(minC=np.empty((H, W), np.uint32), flow=np.empty((2, H, W), np.float64))
    if stages:
        out.update(Centries=np.empty((H, W, D, 3), np.int32), Sp=np.empty((H, W, D), np.uint32))
    g = out.get
    _port().orc_ng(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, P1, P2, C.c_uint(seed), _ptr(out["minC"], _u32p),
                   _ptr(out["flow"], _f64p), _ptr(g("Centries"), _i32p), _ptr(g("Sp"), _u32p))
    return out


def port_pydng(I1, I2, preMv, r, aggSize, subpix, P1, P2, stages=False):
    H, W = I1.shape
    _, mvH, mvW = preMv.shape
    D = 9 * (2 * r + 1) ** 2
    out = dict(minC=np.empty((H, W), np.uint32), flow=np.empty((2, H, W), np.float64))
    if stages:
        out.update(Centries=np.empty((H, W, D, 3), np.int32), Sp=np.empty((H, W, D), np.uint32))
    g = out.get
    _port().orc_pydng(_ptr(I1, _u8p), _ptr(I2, _u8p), W, H, _ptr(preMv, _f64p), mvW, mvH, r, aggSize, subpix, P1, P2,
                      _ptr(out["minC"], _u32p), _ptr(out["flow"], _f64p), _ptr(g("Centries"), _i32p), _ptr(g("Sp"), _u32p))
    return out
