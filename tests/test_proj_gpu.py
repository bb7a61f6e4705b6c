"""GPU end-to-end for the proj/ facade: the sgmof command line (EpiSGM / PydSGM classes, PNG in, KITTI flow PNG
out) against the Python pipeline over the same C ABI, which the other GPU tests pin against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from fsgm_b200 import synth

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu


def _encode(flow):
    f32 = flow.astype(np.float32)
    return np.clip(f32 * np.float32(64.0) + np.float32(32768.0), 0, 65535).astype(np.uint16)


def test_sgmof_pyd_mode(ctx, tmp_path):
    from fsgm_b200 import api, build
    build.build_proj()
    W, H = 200, 120
    fp = synth.flow_pair(W, H, seed=8, umax=6, vmax=3, blocks=2)
    a, b, out = str(tmp_path / "a.png"), str(tmp_path / "b.png"), str(tmp_path / "flow.png")
    cv2.imwrite(a, fp["I1"]); cv2.imwrite(b, np.stack([fp["I2"]] * 3, -1))          # gray and RGB inputs both work
    r = subprocess.run([build.PROJ_BIN, a, b, "-m=1", "-N", "3", f"-o={out}", "-p=2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    mv, _ = ctx.pyramidal_sgm(fp["I1"], fp["I2"], opts=api.pyd_opts(numPyd=3))
    png = cv2.imread(out, cv2.IMREAD_UNCHANGED)[..., ::-1]
    assert png.shape == (H, W, 3) and (png[..., 2] == 1).all()
    assert np.array_equal(png[..., 0], _encode(mv[0])) and np.array_equal(png[..., 1], _encode(mv[1]))
    # ground-truth scoring option: the synthetic flow as a KITTI file
    gt = np.zeros((H, W, 3), np.uint16)
    gt[..., 2], gt[..., 1], gt[..., 0] = _encode(fp["u"]), _encode(fp["v"]), 1                    # OpenCV order B, G, R
    cv2.imwrite(str(tmp_path / "gt.png"), gt)
    r = subprocess.run([build.PROJ_BIN, a, b, "-m=1", "-N=3", f"-o={out}", f"-G={tmp_path / 'gt.png'}"], capture_output=True, text=True)
    assert r.returncode == 0 and "KITTI outliers" in r.stdout
    rate = float(r.stdout.split("KITTI outliers (>3 px and >5 %):")[1].split("%")[0])
    err = np.hypot(mv[0] - fp["u"], mv[1] - fp["v"])
    mag = np.hypot(fp["u"], fp["v"])
    want = 100.0 * ((err > 3) & (err > 0.05 * mag)).mean()
    assert abs(rate - want) < 0.5                         # the file quantises to 1/64 px


def test_sgmof_epi_mode(ctx, tmp_path):
    from fsgm_b200 import api, build
    build.build_proj()
    W, H, D = 160, 96, 64
    p = synth.epipolar_pair(W, H, D, seed=5)
    cam = synth.epipolar_camera(W, H, seed=6, rot_deg=0.05)
    a, b, out = str(tmp_path / "a.png"), str(tmp_path / "b.png"), str(tmp_path / "flow.png")
    cv2.imwrite(a, p["I1"]); cv2.imwrite(b, p["I2"])
    (tmp_path / "calib.txt").write_text("P0: 7.070912e+02 0 6.018873e+02 0 0 7.070912e+02 1.831104e+02 0 0 0 1 0\n")
    geo = list(cam["F"]) + list(cam["H"]) + list(cam["epi"]) + [cam["direction"]]
    (tmp_path / "geo.txt").write_text(" ".join(repr(float(v)) for v in geo))
    args = [build.PROJ_BIN, a, b, "-m=0", f"-c={tmp_path / 'calib.txt'}", f"-g={tmp_path / 'geo.txt'}", f"-o={out}", "-d"]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    flow, _ = ctx.epipolar_sgm_of_batch(p["I1"][None], p["I2"][None], cam["F"], cam["H"], cam["epi"], [cam["direction"]], 64, 0.3, 6, 64,
                                        opts=api.epi_opts(paths=8))
    png = cv2.imread(out, cv2.IMREAD_UNCHANGED)[..., ::-1]
    assert np.array_equal(png[..., 0], _encode(flow[0, 0])) and np.array_equal(png[..., 1], _encode(flow[0, 1]))
    # without geometry the epipolar mode refuses (the reference's C++ never estimates it either)
    r = subprocess.run(args[:5] + [f"-o={out}"], capture_output=True, text=True)
    assert r.returncode == 1 and "needs -g" in r.stdout
