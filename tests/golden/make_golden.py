"""Regenerates tests/golden/*.npz from the reference's own C++ (oracle/_ref, built from /root/reference).

Run in the authoring container only:   python tests/golden/make_golden.py
The fixtures hold inputs AND the reference's outputs (gateway outputs plus the intermediate buffers fished out
of the MEX shim's allocation log), so the parity tests on the GPU box never need /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from fsgm_b200 import synth          # noqa: E402
from oracle import pyoracle as po    # noqa: E402


ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]     # fixture names to (re)write; default all


def save(name, **arrs):
    if ONLY and name not in ONLY:
        return
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    po.build(port=False, ref=True)
    rng = np.random.default_rng(2024)
    # ---- epipolar, 4 paths (as shipped) and 8 paths; one case outside the no-wrap domain -----------------
    for name, (W, H, D, P1, P2, paths) in {
        "epi_p8": (56, 36, 32, 6, 64, 8), "epi_p4": (56, 36, 32, 6, 64, 4),
        "epi_odd": (41, 23, 13, 6, 64, 8), "epi_wrap": (40, 24, 16, 100, 200, 8),
    }.items():
        p = synth.epipolar_pair(W, H, D, seed=7)
        r = po.ref_epi(p["I1"], p["I2"], D, p["vMax"], p["Pd0"], p["dirn"], p["O"], P1, P2, paths=paths)
        save(name, I1=p["I1"], I2=p["I2"], Pd0=p["Pd0"], dirn=p["dirn"], O=p["O"], vMax=p["vMax"], D=D, P1=P1, P2=P2,
             paths=paths, cen1=r["cen1"], cen2=r["cen2"], Craw=r["Craw"], C=r["C"], Sp=r["Sp"].astype(np.uint16),
             bestD=r["bestD"], minC=r["minC"])
    # ---- forward/backward check: the 8-path build with the call at calc_cost_sgm.cpp:589-590 re-enabled, and the function
    #      called directly on a random x256 label map with a loose threshold -------------------------------------------------
    p = synth.epipolar_pair(56, 36, 32, seed=7)
    r = po.ref_epi_fb(p["I1"], p["I2"], 32, p["vMax"], p["Pd0"], p["dirn"], p["O"], 6, 64)
    D1 = np.random.default_rng(77).integers(0, 32 * 256, (36, 56)).astype(np.uint32)
    conf_d, D2_d = po.ref_fb_check(D1, p["Pd0"], p["dirn"], p["O"], p["vMax"], 33, thr=600)
    save("epi_fb", I1=p["I1"], I2=p["I2"], Pd0=p["Pd0"], dirn=p["dirn"], O=p["O"], vMax=p["vMax"], D=32, P1=6, P2=64,
         bestD=r["bestD"], minC=r["minC"], conf=r["conf"], bestD2=r["bestD2"], D1=D1, thr_d=600, conf_d=conf_d, bestD2_d=D2_d,
         disp_d=po.ref_vz_to_disp(D1, p["O"], p["vMax"], 33))
    # ---- pyramidal: integer prior, fractional prior with adaptive P2, single pass without diagonals ---------
    fp = synth.flow_pair(48, 32, seed=5, umax=3, vmax=2)
    for name, (rx, ry, sub, diag, passes, adp, kind) in {
        "pyd_a": (2, 2, 1, 1, 2, 0, "int"), "pyd_b": (3, 2, 1, 1, 2, 1, "frac"), "pyd_c": (2, 3, 0, 0, 1, 0, "int"),
    }.items():
        mv = np.zeros((2, 36, 52))
        mv[:] = rng.integers(-3, 4, mv.shape) if kind == "int" else rng.normal(0, 2, mv.shape)
        r = po.ref_pyd(fp["I1"], fp["I2"], mv, rx, ry, 2, sub, 6, 32, diag, passes, adp)
        save(name, I1=fp["I1"], I2=fp["I2"], preMv=mv, rx=rx, ry=ry, agg=2, sub=sub, P1=6, P2=32, diag=diag, passes=passes,
             adaptive=adp, C=r["C"], Sp=r["Sp"].astype(np.uint16), bestD=r["bestD"], minC=r["minC"], mvSub=r["mvSub"])
    # ---- pyramid driver (pyramidal_sgm.m): per-level solver = the reference build; impyramid restated (oracle/pyramid_oracle.py) ----
    from oracle import pyramid_oracle as pyo
    fpp = synth.flow_pair(90, 58, seed=21, umax=9, vmax=5)
    mv, mc, lv = pyo.pyramidal_sgm(fpp["I1"], fpp["I2"], lambda *a: po.ref_pyd(*a, stages=False), numPyd=3, ver=2, hor=3)
    save("pyramid_a", I0=fpp["I1"], I1=fpp["I2"], numPyd=3, ver=2, hor=3, mv=mv, minC=mc,
         **{f"mv_l{i}": m for i, m in enumerate(lv)}, **{f"img_l{i}": m for i, m in enumerate(pyo.pyramid(fpp["I1"], 3))})
    # ---- neighbour guided (glibc rand(), srand(1)) ---------------------------------------------------------
    fq = synth.flow_pair(28, 20, seed=9, umax=3, vmax=2)
    r = po.ref_ng(fq["I1"], fq["I2"], 6, 32, seed=1, stages=True)
    save("ng_a", I1=fq["I1"], I2=fq["I2"], P1=6, P2=32, seed=1, Centries=r["Centries"].astype(np.int16),
         Sp=r["Sp"].astype(np.uint16), minC=r["minC"], flow=r["flow"])
    mv = rng.normal(0, 2, (2, 24, 30))
    for name, sub in (("pydng_a", 0), ("pydng_b", 1)):
        r = po.ref_pydng(fq["I1"], fq["I2"], mv, 1, 5, sub, 6, 32, stages=True)
        save(name, I1=fq["I1"], I2=fq["I2"], preMv=mv, r=1, aggSize=5, sub=sub, P1=6, P2=32,
             Centries=r["Centries"].astype(np.int16), Sp=r["Sp"].astype(np.uint16), minC=r["minC"], flow=r["flow"])


if __name__ == "__main__":
    main()
