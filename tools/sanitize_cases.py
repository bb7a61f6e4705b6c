#!/usr/bin/env python
"""Small cases, one per kernel family, for compute-sanitizer (memcheck / racecheck on every kernel).

  compute-sanitizer --tool memcheck  python tools/sanitize_cases.py [family ...]
  compute-sanitizer --tool racecheck python tools/sanitize_cases.py [family ...]

Every case goes through the C ABI exactly like the parity tests and is small enough for the sanitizer's 10-100x slowdown.
Families: epi_generic, epi_cluster2, epi_cluster9, epi_wrap, pyd, pyramid, ng, pydng, geometry, fbcheck, stage."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                            # noqa: E402
from fsgm_b200 import api, synth                        # noqa: E402


def stack(ps, k):
    return np.ascontiguousarray(np.stack([p[k] for p in ps]))


def epi(ctx, W, H, D, n, cluster, P1=6, P2=64, paths=8):
    ps = [synth.epipolar_pair(W, H, D, seed=40 + i) for i in range(n)]
    ctx.tune(1, cluster)
    o = api.epi_opts(paths=paths)
    b, m = ctx.calc_cost_sgm_batch(stack(ps, "I1"), stack(ps, "I2"), D, 0.3, stack(ps, "Pd0"), stack(ps, "dirn"), stack(ps, "O"),
                                   P1, P2, opts=o)
    ctx.tune(1, 0)
    return int(b.sum() % 1000003), int(m.sum() % 1000003)


def main():
    want = set(sys.argv[1:])
    ctx = api.Context(0)
    fam = {}
    fam["epi_generic"] = lambda: epi(ctx, 70, 24, 40, 1, -1)                 # LM_VECPAD labels, generic sweeps + WTA kernel
    fam["epi_cluster2"] = lambda: epi(ctx, 64, 20, 64, 2, 2)                 # fused cost, hsweep_tma, vsweep cluster of 2, finalize
    fam["epi_cluster9"] = lambda: epi(ctx, 99, 16, 64, 2, 9)                 # non-portable cluster size, 11 columns per CTA
    fam["epi_cluster4_d256_4paths"] = lambda: epi(ctx, 80, 12, 256, 1, 4, paths=4)
    fam["epi_wrap"] = lambda: epi(ctx, 40, 16, 24, 1, 0, P1=200, P2=250)     # explicit mod-256 sweep kernel
    fp = synth.flow_pair(48, 30, seed=3, umax=3, vmax=2)
    mv = np.round(np.random.default_rng(1).normal(0, 1.5, (2, 30, 48)))
    fam["pyd"] = lambda: [a.sum() for a in ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv, 5, 5, 2, 1, 6, 32, 1, 2, 0)][:2]
    fam["pyd_frac_prior"] = lambda: [a.sum() for a in ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv * 0.37, 3, 2, 1, 1, 6, 32, 1, 2, 1)][:2]
    fam["pyramid"] = lambda: [a.sum() for a in ctx.pyramidal_sgm(fp["I1"], fp["I2"], opts=api.pyd_opts(numPyd=3, ver=2, hor=2))]
    fam["ng"] = lambda: [a.sum() for a in ctx.calc_cost_sgm_ng(fp["I1"][:12, :24].copy(), fp["I2"][:12, :24].copy(), P1=6, P2=32, seed=1)]
    fam["pydng"] = lambda: [a.sum() for a in ctx.calc_pyd_cost_sgm_ng(fp["I1"][:16, :24].copy(), fp["I2"][:16, :24].copy(),
                                                                      mv[:, :16, :24].copy(), 1, 5, 1, 6, 32)]

    def geometry():
        cam = synth.epipolar_camera(64, 20, seed=2, rot_deg=0.05)
        p = synth.epipolar_pair(64, 20, 64, seed=2)
        f, m = ctx.epipolar_sgm_of_batch(p["I1"][None], p["I2"][None], [cam["F"]], [cam["H"]], [cam["epi"]], [cam["direction"]],
                                         64, 0.3, 6, 64, opts=api.epi_opts(paths=8), f32=True)
        return float(f.sum()), int(m.sum())
    fam["geometry"] = geometry

    def fbcheck():
        p = synth.epipolar_pair(64, 20, 64, seed=5)
        r = ctx.calc_cost_sgm(p["I1"], p["I2"], 64, 0.3, p["Pd0"], p["dirn"], p["O"], 6, 64, opts=api.epi_opts(paths=8, fb_check=1))
        return [int(a.sum()) for a in r]
    fam["fbcheck"] = fbcheck

    def stage():
        rng = np.random.default_rng(3)
        Cv = torch.from_numpy(rng.integers(0, 200, (1, 12, 21, 40), dtype=np.uint8)).cuda()
        I1 = torch.from_numpy(rng.integers(0, 256, (1, 12, 21), dtype=np.uint8)).cuda()
        b = torch.empty((1, 12, 21), dtype=torch.int32, device="cuda"); m = torch.empty_like(b)
        O = torch.ones((1, 12, 21), dtype=torch.float64, device="cuda")
        ctx.epi_aggregate_dev(Cv, I1, 6, 64, O, 0.3, b, m, opts=api.epi_opts(paths=8, vz_to_disp=0))
        return int(b.sum())
    fam["stage"] = stage

    for name, fn in fam.items():
        if want and name not in want:
            continue
        r = fn()
        torch.cuda.synchronize()
        print(f"case {name}: done, checksum {r}", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
