"""The reference-side binding, compiled and exercised: the four replacement MEX gateways under integration/mex/
are built against the same mex.h shim the reference's own gateways are built with (oracle/Makefile), called through the
reference's real gateway signature `mexFunction(nlhs, plhs, nrhs, prhs)` and compared with oracle/_ref / the restatement."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from fsgm_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "_mexbuild")
STUBS = {                                           # name -> (stub source, harness variant, extra flags)
    "epi4": ("calc_cost_sgm.cpp", "STUB_EPI", []),
    "epi8": ("calc_cost_sgm.cpp", "STUB_EPI", ["-DFSGM_MEX_PATHS=8"]),
    "pyd": ("calc_pyd_cost_sgm.cpp", "STUB_PYD", []),
    "ng": ("calc_cost_sgm_ng.cpp", "STUB_NG", []),
    "pydng": ("calc_pyd_cost_sgm_ng.cpp", "STUB_PYDNG", []),
}


def build_stub(name):
    src, variant, extra = STUBS[name]
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, f"libmexstub_{name}.so")
    libdir = os.path.join(ROOT, "fsgm_b200")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Werror", f"-D{variant}", *extra,
           "-I" + os.path.join(ROOT, "oracle", "mex_shim"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "integration", "mex", src), os.path.join(ROOT, "tests", "mex_stub_driver.cpp"),
           "-L" + libdir, "-lfsgm", "-Wl,-rpath," + libdir, "-o", so]
    subprocess.run(cmd, check=True)
    return so


@pytest.mark.parametrize("name", sorted(STUBS))
def test_stubs_compile_link_and_export_the_gateway(name):
    """CPU: every stub builds warning-clean against mex.h + include/fsgm.h, links libfsgm.so and exports mexFunction."""
    so = build_stub(name)
    syms = subprocess.run(["nm", "-D", "--defined-only", so], check=True, capture_output=True, text=True).stdout
    assert "mexFunction" in syms
    needed = subprocess.run(["readelf", "-d", so], check=True, capture_output=True, text=True).stdout
    assert "libfsgm.so" in needed


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def _load(name):
    lib = C.CDLL(build_stub(name))
    lib.stub_last_error.restype = C.c_char_p
    return lib


@pytest.mark.gpu
@pytest.mark.parametrize("paths", [4, 8])
def test_mex_stub_calc_cost_sgm(ctx, oracle, paths):
    W, H, D = 96, 60, 64
    p = synth.epipolar_pair(W, H, D, seed=31)
    lib = _load(f"epi{paths}")
    bestD, minC = np.empty((H, W), np.uint32), np.empty((H, W), np.uint32)
    conf, bestD2 = np.full((H, W), 7, np.uint8), np.full((H, W), 7, np.uint32)
    rc = lib.stub_epi(_p(p["I1"], C.c_uint8), _p(p["I2"], C.c_uint8), W, H, D, C.c_double(p["vMax"]), _p(p["Pd0"], C.c_double),
                      _p(p["dirn"], C.c_double), _p(p["O"], C.c_double), 6, 64, _p(bestD, C.c_uint32), _p(minC, C.c_uint32),
                      _p(conf, C.c_uint8), _p(bestD2, C.c_uint32))
    assert rc == 0, lib.stub_last_error()
    f = oracle.ref_epi if oracle.have_ref(f"epi{paths}") else oracle.port_epi
    ref = f(p["I1"], p["I2"], D, p["vMax"], p["Pd0"], p["dirn"], p["O"], 6, 64, paths=paths)
    assert np.array_equal(minC, ref["minC"])
    a, b = bestD.copy(), ref["bestD"].copy()
    if ref["Sp"][-1, -1].argmin() == D - 1:
        a[-1, -1] = b[-1, -1] = 0
    assert np.array_equal(a, b)
    assert not conf.any() and not bestD2.any()          # outputs 3 and 4 stay zero as shipped (calc_cost_sgm.cpp:571-572, :589-590)
    lib.stub_shutdown()


@pytest.mark.gpu
def test_mex_stub_calc_pyd_cost_sgm(ctx, oracle):
    W, H = 70, 44
    fp = synth.flow_pair(W, H, seed=5, umax=3, vmax=2)
    rng = np.random.default_rng(2)
    mvW, mvH = W + 2, H + 1                              # preMv has its own stride (calc_pyd_cost_sgm.cpp:388, :493-494)
    pre = np.ascontiguousarray(rng.uniform(-1.5, 1.5, (2, mvH, mvW)))
    lib = _load("pyd")
    bestD, minC, mvSub = np.empty((H, W), np.uint32), np.empty((H, W), np.uint32), np.empty((2, H, W), np.float64)
    rc = lib.stub_pyd(_p(fp["I1"], C.c_uint8), _p(fp["I2"], C.c_uint8), W, H, _p(pre, C.c_double), mvW, mvH, 3, 2, 2, 1, 6, 32, 1, 2, 0,
                      _p(bestD, C.c_uint32), _p(minC, C.c_uint32), _p(mvSub, C.c_double))
    assert rc == 0, lib.stub_last_error()
    f = oracle.ref_pyd if oracle.have_ref("pyd") else oracle.port_pyd
    ref = f(fp["I1"], fp["I2"], pre, 3, 2, 2, 1, 6, 32, 1, 2, 0, stages=False)
    assert np.array_equal(bestD, ref["bestD"]) and np.array_equal(minC, ref["minC"])
    assert np.array_equal(mvSub, ref["mvSub"])
    lib.stub_shutdown()


@pytest.mark.gpu
def test_mex_stub_calc_cost_sgm_ng_continuing_rand_stream(ctx, oracle):
    """The stub draws its hints from the process's rand() like the reference (calc_cost_sgm_ng.cpp:148-149): the first call after
    srand(1) equals the pinned oracle, and a second call continues the stream instead of restarting it."""
    from fsgm_b200 import api
    W, H = 36, 24
    fp = synth.flow_pair(W, H, seed=8, umax=3, vmax=2)
    pre = np.zeros((2, H, W))
    lib = _load("ng")
    minC, flow = np.empty((H, W), np.uint32), np.empty((2, H, W), np.float64)
    args = (_p(fp["I1"], C.c_uint8), _p(fp["I2"], C.c_uint8), W, H, _p(pre, C.c_double), W, H, C.c_double(1), C.c_double(5), 0, 6, 32)
    rc = lib.stub_ng(*args, 1, _p(minC, C.c_uint32), _p(flow, C.c_double))
    assert rc == 0, lib.stub_last_error()
    f = oracle.ref_ng if oracle.have_ref("ng") else oracle.port_ng
    ref = f(fp["I1"], fp["I2"], 6, 32, seed=1)
    assert np.array_equal(minC, ref["minC"]) and np.array_equal(flow, ref["flow"])
    rc = lib.stub_ng(*args, -1, _p(minC, C.c_uint32), _p(flow, C.c_double))
    assert rc == 0, lib.stub_last_error()
    second = api.glibc_rand(1, 16 * W * H)[8 * W * H:].copy()
    m2, f2 = ctx.calc_cost_sgm_ng(fp["I1"], fp["I2"], P1=6, P2=32, rand_stream=second)
    assert np.array_equal(minC, m2) and np.array_equal(flow, f2)
    assert not np.array_equal(flow, ref["flow"])
    lib.stub_shutdown()


@pytest.mark.gpu
def test_mex_stub_calc_pyd_cost_sgm_ng(ctx, oracle):
    W, H = 40, 28
    fp = synth.flow_pair(W, H, seed=9, umax=3, vmax=2)
    rng = np.random.default_rng(4)
    pre = np.ascontiguousarray(rng.uniform(-2, 2, (2, H, W)))
    lib = _load("pydng")
    minC, flow = np.empty((H, W), np.uint32), np.empty((2, H, W), np.float64)
    rc = lib.stub_ng(_p(fp["I1"], C.c_uint8), _p(fp["I2"], C.c_uint8), W, H, _p(pre, C.c_double), W, H, C.c_double(1), C.c_double(5), 1,
                     6, 32, 1, _p(minC, C.c_uint32), _p(flow, C.c_double))
    assert rc == 0, lib.stub_last_error()
    f = oracle.ref_pydng if oracle.have_ref("pydng") else oracle.port_pydng
    ref = f(fp["I1"], fp["I2"], pre, 1, 5, 1, 6, 32)
    assert np.array_equal(minC, ref["minC"]) and np.array_equal(flow, ref["flow"])
    lib.stub_shutdown()
