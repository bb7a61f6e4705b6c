"""CPU suite for the proj/ facade's host-side I/O: PNG codec against OpenCV's, the KITTI flow encoding of
proj/src/utils.cpp:3-73, the calibration reader of :129-169, and the command line's argument handling."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def proj():
    from fsgm_b200 import build
    build.build()
    build.build_proj()
    return C.CDLL(build.PROJ_LIB)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _read(proj, path):
    r, c, ch, bd = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert proj.fsgm_proj_png_info(path.encode(), C.byref(r), C.byref(c), C.byref(ch), C.byref(bd)) == 0
    s = np.empty((r.value, c.value, ch.value), np.uint16)
    assert proj.fsgm_proj_png_read(path.encode(), _p(s, C.c_uint16)) == 0
    return s, bd.value


@pytest.mark.parametrize("shape,dtype", [((37, 53), np.uint8), ((20, 31, 3), np.uint8), ((16, 9, 3), np.uint16), ((5, 7), np.uint16),
                                         ((11, 13, 4), np.uint8)])
def test_png_codec_against_opencv(proj, tmp_path, shape, dtype):
    rng = np.random.default_rng(sum(shape))
    img = rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
    img[..., 0] = np.sort(img[..., 0], axis=-1) if img.ndim == 3 else img[..., 0]       # some structure: exercises cv2's row filters
    a = str(tmp_path / "cv.png")
    assert cv2.imwrite(a, img)                                                        # OpenCV picks filters per row
    got, depth = _read(proj, a)
    want = img if img.ndim == 2 else (img[..., ::-1] if shape[2] == 3 else img[..., [2, 1, 0, 3]])      # file order is RGB(A)
    assert depth == 8 * img.itemsize and np.array_equal(got.reshape(want.shape), want)
    b = str(tmp_path / "ours.png")
    ch = 1 if img.ndim == 2 else shape[2]
    s16 = np.ascontiguousarray(want.astype(np.uint16))
    assert proj.fsgm_proj_png_write(b.encode(), shape[0], shape[1], ch, 8 * img.itemsize, _p(s16, C.c_uint16)) == 0
    back = cv2.imread(b, cv2.IMREAD_UNCHANGED)
    assert back.dtype == dtype and np.array_equal(back, img)


def test_imread_gray_matches_rgb2gray(proj, tmp_path):
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (24, 40, 3)).astype(np.uint8)
    path = str(tmp_path / "c.png")
    cv2.imwrite(path, rgb[..., ::-1])
    g = np.empty((24, 40), np.uint8)
    assert proj.fsgm_proj_imread_gray(path.encode(), _p(g, C.c_uint8)) == 0
    want = np.floor(0.298936021293775 * rgb[..., 0] + 0.587043074451121 * rgb[..., 1] + 0.114020904255103 * rgb[..., 2] + 0.5)
    assert np.array_equal(g, want.astype(np.uint8))


def test_kitti_flow_encoding(proj, tmp_path):
    """(u*64 + 32768) in float, truncated, clamped to 0..65535; file channels R = u, G = v, B = valid (utils.cpp:55-66)"""
    H, W = 6, 8
    uv = np.zeros((H, W, 2), np.float32)
    valid = np.ones((H, W), np.uint8)
    uv[0, 0] = (0.0, 0.0); uv[0, 1] = (1.0, -1.0); uv[0, 2] = (0.0078125, 511.99); uv[0, 3] = (-512.5, 600.0); uv[0, 4] = (3.3, -7.77)
    valid[1, :] = 0; uv[1, :] = 5.0
    path = str(tmp_path / "f.png")
    assert proj.fsgm_proj_flow_write(path.encode(), H, W, _p(uv, C.c_float), _p(valid, C.c_uint8)) == 0
    raw = cv2.imread(path, cv2.IMREAD_UNCHANGED)[..., ::-1]                       # -> R, G, B
    assert raw.dtype == np.uint16
    enc = lambda x: np.uint16(max(min(np.float32(x) * np.float32(64.0) + np.float32(32768.0), np.float32(65535.0)), np.float32(0.0)))
    assert tuple(raw[0, 0]) == (32768, 32768, 1) and tuple(raw[0, 1]) == (32832, 32704, 1)
    assert tuple(raw[0, 2]) == (32768, enc(511.99), 1) and tuple(raw[0, 3]) == (0, 65535, 1)
    assert tuple(raw[0, 4]) == (enc(3.3), enc(-7.77), 1)
    assert not raw[1].any()
    uv2 = np.empty_like(uv); v2 = np.empty_like(valid)
    assert proj.fsgm_proj_flow_read(path.encode(), _p(uv2, C.c_float), _p(v2, C.c_uint8)) == 0
    assert np.array_equal(v2, valid)
    dec = (raw[..., :2].astype(np.float32) - np.float32(32768.0)) / np.float32(64.0)
    assert np.array_equal(uv2[valid == 1], dec[valid == 1]) and not uv2[valid == 0].any()
    assert np.abs(uv2[0, 4] - uv[0, 4]).max() <= 1 / 64


def test_calib_reader_2012_and_2015(proj, tmp_path):
    P0 = [7.070912e+02, 0, 6.018873e+02, 0, 0, 7.070912e+02, 1.831104e+02, 0, 0, 0, 1, 0]
    f12 = tmp_path / "c12.txt"
    f12.write_text("P0: " + " ".join(f"{v:e}" for v in P0) + "\nP1: " + " ".join("1" for _ in range(12)) + "\n")
    out = np.empty(12, np.float32)
    assert proj.fsgm_proj_read_calib(str(f12).encode(), 0, _p(out, C.c_float)) == 0
    assert np.allclose(out, np.float32(P0))
    f15 = tmp_path / "c15.txt"
    f15.write_text("".join(f"line{i}: 0 0 0\n" for i in range(9)) + "P_rect_00: " + " ".join(f"{v:e}" for v in P0) + "\n")
    assert proj.fsgm_proj_read_calib(str(f15).encode(), 1, _p(out, C.c_float)) == 0
    assert np.allclose(out, np.float32(P0))
    assert proj.fsgm_proj_read_calib(str(tmp_path / "missing.txt").encode(), 0, _p(out, C.c_float)) != 0


def test_reference_example_files_decode_like_opencv(proj):
    ex = "/root/reference/proj/example"
    if not os.path.isdir(ex):
        pytest.skip("reference example files are only present in the authoring container")
    for name in ("000000_10.png", "000000_10_gtFlow.png"):
        got, _ = _read(proj, os.path.join(ex, name))
        want = cv2.imread(os.path.join(ex, name), cv2.IMREAD_UNCHANGED)[..., ::-1]
        assert np.array_equal(got, want)
    out = np.empty(12, np.float32)
    assert proj.fsgm_proj_read_calib(os.path.join(ex, "000000.txt").encode(), 0, _p(out, C.c_float)) == 0
    assert abs(out[0] - 707.0912) < 1e-3 and abs(out[2] - 601.8873) < 1e-3


def test_cli_usage_and_errors(proj, tmp_path):
    from fsgm_b200 import build
    r = subprocess.run([build.PROJ_BIN, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "SGM OF v0.0.1" in r.stdout and "-N, --pydNum" in r.stdout
    r = subprocess.run([build.PROJ_BIN], capture_output=True, text=True)
    assert r.returncode == 1
    r = subprocess.run([build.PROJ_BIN, str(tmp_path / "a.png"), str(tmp_path / "b.png"), "-m=1"], capture_output=True, text=True)
    assert r.returncode == 1 and "Open image failed" in r.stdout
    cv2.imwrite(str(tmp_path / "a.png"), np.zeros((8, 9), np.uint8)); cv2.imwrite(str(tmp_path / "b.png"), np.zeros((9, 9), np.uint8))
    r = subprocess.run([build.PROJ_BIN, str(tmp_path / "a.png"), str(tmp_path / "b.png"), "-m=1"], capture_output=True, text=True)
    assert r.returncode == 1 and "Size of image1/2 must match" in r.stdout
