#!/usr/bin/env python
"""Secondary measurements (not the bench.py contract): throughput of the three other gateways on one B200 —
calc_pyd_cost_sgm (configs[2]), calc_cost_sgm_ng and calc_pyd_cost_sgm_ng (configs[3]) — with the reference CPU time of
the same call beside it on a bounded sample.  These paths are integer-issue / latency bound, not HBM bound
(SURVEY.md §8d), so they are reported as pairs/s and label evaluations/s only.

  python bench_variants.py [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    from fsgm_b200 import api, synth
    from oracle import pyoracle as po
    ctx = api.Context(0)
    ctx.use_torch_stream()
    use_ref = po.have_ref("pyd")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def timed(fn, reps):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {}
    W, H = 1242, 375
    fp = synth.flow_pair(W, H, seed=1, umax=20, vmax=10)
    # ---- pyramidal: 3 levels, r = 5 (reference default, D = 121), 8 paths, 2 passes ------------------------------
    lv = [(fp["I1"], fp["I2"])]
    for _ in range(2):
        lv.append((synth.reduce2(lv[-1][0]), synth.reduce2(lv[-1][1])))
    n = 4 if args.quick else 8
    tot_ms, evals = 0.0, 0
    for lvl, (a, b) in enumerate(lv):
        h, w = a.shape
        I1 = t(np.stack([a] * n)); I2 = t(np.stack([b] * n))
        mv = torch.zeros((n, 2, h, w), dtype=torch.float64, device="cuda")
        bD = torch.empty((n, h, w), dtype=torch.int32, device="cuda"); mC = torch.empty_like(bD)
        ms_ = torch.empty((n, 2, h, w), dtype=torch.float64, device="cuda")
        ms = timed(lambda: ctx.calc_pyd_cost_sgm_dev(I1, I2, mv, 5, 5, 2, int(lvl == 0), 6, 32, 1, 2, 0, bD, mC, ms_), 3)
        out[f"pyd_level{lvl}_{w}x{h}_ms_per_pair"] = ms / n
        tot_ms += ms / n; evals += w * h * 121
    out["pyd_3level_pairs_per_s"] = 1e3 / tot_ms
    out["pyd_3level_gde_per_s"] = evals / (tot_ms * 1e-3) / 1e9
    # ---- N1: the whole pyramid loop on the device (pyramidal_sgm.m), 3 and 5 levels, r = 5 ---------------------------------
    n = 4 if args.quick else 8
    I0 = t(np.stack([fp["I1"]] * n)); I1b = t(np.stack([fp["I2"]] * n))
    mvo = torch.empty((n, 2, H, W), dtype=torch.float64, device="cuda")
    mCo = torch.empty((n, H, W), dtype=torch.int32, device="cuda")
    for L in (3, 5):
        o = api.pyd_opts(numPyd=L)
        ms = timed(lambda: ctx.pyramidal_sgm_dev(I0, I1b, mvo, mCo, opts=o), 3)
        out[f"pyramid_driver_{L}level_pairs_per_s"] = n / (ms * 1e-3)
    del I0, I1b, mvo, mCo
    # ---- N2: F/H/epipole -> flow through the host-image call (2 B/px up, 20 B/px back) vs gateway 1 (42 up, 8 back) ---------
    if not args.quick:
        npairs, D = 60, 256
        ep = synth.epipolar_pair(W, H, D, seed=1)
        cam = synth.epipolar_camera(W, H, seed=2, rot_deg=0.05)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        hI0 = pin(np.stack([ep["I1"]] * npairs)); hI1 = pin(np.stack([ep["I2"]] * npairs))
        hflow = torch.empty((npairs, 2, H, W), dtype=torch.float64).pin_memory()
        hmin = torch.empty((npairs, H, W), dtype=torch.int32).pin_memory()
        Fs, Hs, es, ds = [cam["F"]] * npairs, [cam["H"]] * npairs, [cam["epi"]] * npairs, [cam["direction"]] * npairs
        outs = (hflow.numpy(), hmin.numpy().view(np.uint32))
        o8 = api.epi_opts(paths=8)
        call = lambda: ctx.epipolar_sgm_of_batch(hI0.numpy(), hI1.numpy(), Fs, Hs, es, ds, D, 0.3, 6, 64, opts=o8, out=outs, asynchronous=True)
        for _ in range(3):
            call()
        ctx.synchronize()
        t0 = time.perf_counter()
        reps = 8
        for _ in range(reps):
            call()
        ctx.synchronize()
        dt = time.perf_counter() - t0
        out["epipolar_sgm_of_e2e_pairs_per_s"] = npairs * reps / dt
        out["epipolar_sgm_of_h2d_bytes_per_pair"] = 2 * W * H
        out["epipolar_sgm_of_d2h_bytes_per_pair"] = 20 * W * H
        del hI0, hI1, hflow, hmin
    # ---- pyd_ng r = 1 (D = 81 candidates), aggSize 5 -----------------------------------------------------------
    n = 4 if args.quick else 8
    I1 = t(np.stack([fp["I1"]] * n)); I2 = t(np.stack([fp["I2"]] * n))
    mv = torch.zeros((n, 2, H, W), dtype=torch.float64, device="cuda")
    mC = torch.empty((n, H, W), dtype=torch.int32, device="cuda")
    fl = torch.empty((n, 2, H, W), dtype=torch.float64, device="cuda")
    ms = timed(lambda: ctx.calc_pyd_cost_sgm_ng_dev(I1, I2, mv, 1, 5, 1, 6, 32, mC, fl), 2)
    out["pydng_r1_pairs_per_s"] = n / (ms * 1e-3)
    out["pydng_r1_candidate_tests_per_s"] = n * W * H * 4 * 81 * 81 / (ms * 1e-3)
    # ---- ng: one CTA per pair, so throughput needs a batch of >= 148 pairs; measure a reduced-height strip -----------
    hs = 24 if args.quick else 48
    n = 148
    sp = synth.flow_pair(W, hs, seed=2, umax=20, vmax=4)
    I1 = t(np.stack([sp["I1"]] * n)); I2 = t(np.stack([sp["I2"]] * n))
    mC = torch.empty((n, hs, W), dtype=torch.int32, device="cuda")
    fl = torch.empty((n, 2, hs, W), dtype=torch.float64, device="cuda")
    ms = timed(lambda: ctx.calc_cost_sgm_ng_dev(I1, I2, 6, 32, mC, fl, seeds=list(range(1, n + 1))), 1)
    px_s = n * W * hs / (ms * 1e-3)
    out["ng_pixels_per_s_148_pairs"] = px_s
    out["ng_kitti_pairs_per_s_extrapolated"] = px_s / (W * H)
    out["ng_us_per_pixel_per_cta"] = ms * 1e3 / (W * hs)
    out["ng_sample"] = f"148 pairs of {W}x{hs} (one CTA each), per-pixel cost is size-independent"
    # ---- CPU reference on bounded samples -----------------------------------------------------------------
    if not args.no_cpu:
        a, b = lv[2]
        h, w = a.shape
        f = po.ref_pyd if use_ref else po.port_pyd
        t0 = time.perf_counter(); f(a, b, np.zeros((2, h, w)), 5, 5, 2, 0, 6, 32, 1, 2, 0, stages=False); dt = time.perf_counter() - t0
        out["cpu_pyd_level2_s"] = dt
        out["cpu_pyd_3level_pairs_per_s_extrapolated"] = 1.0 / (dt * (1 + 4 + 16))
        s2 = synth.flow_pair(W // 4, 32, seed=3, umax=8, vmax=4)
        f = po.ref_ng if use_ref else po.port_ng
        t0 = time.perf_counter(); f(s2["I1"], s2["I2"], 6, 32, seed=1); dt = time.perf_counter() - t0
        out["cpu_ng_pixels_per_s"] = (W // 4) * 32 / dt
        f = po.ref_pydng if use_ref else po.port_pydng
        t0 = time.perf_counter(); f(s2["I1"], s2["I2"], np.zeros((2, 32, W // 4)), 1, 5, 1, 6, 32); dt = time.perf_counter() - t0
        out["cpu_pydng_r1_pixels_per_s"] = (W // 4) * 32 / dt
        out["cpu_kind"] = "reference" if use_ref else "port"
    print(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
