#!/usr/bin/env python
"""Randomised sweep of calc_pyd_cost_sgm (gateway, whole path) against the CPU oracle: random shapes, search windows, aggregation
radii and PIECEWISE-CONSTANT integer priors (blocks of random size: uniform interiors take the separable cost kernel, block borders
and image borders the list kernel's clean and checked paths), now and then a fractional or wild prior.  Not a test: prints and exits
non-zero on the first mismatch."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fsgm_b200 import api, synth
from oracle import pyoracle as po

ctx = api.Context(0)
rng = np.random.default_rng(int(os.environ.get("SEED", "11")))
n_cases = int(os.environ.get("CASES", "30"))
oracle = po.port_pyd            # the C restatement (pinned against the reference build in tests/test_oracle_cpu.py); runs on the GPU box
for it in range(n_cases):
    W = int(rng.integers(20, 150)); H = int(rng.integers(12, 90))
    rx = int(rng.integers(1, 6)); ry = int(rng.integers(1, 6)); agg = int(rng.choice([1, 2, 2]))
    fp = synth.flow_pair(W, H, seed=int(rng.integers(1 << 30)), umax=max(1, rx - 1), vmax=max(1, ry - 1))
    bs = int(rng.choice([2, 4, 8, 16, 32]))
    kind = rng.choice(["blocks", "blocks", "blocks", "frac", "mixed"])
    by, bx = (H + bs - 1) // bs, (W + bs - 1) // bs
    mv = np.zeros((2, H, W))
    for c in range(2):
        coarse = rng.integers(-6, 7, (by, bx)).astype(np.float64) * (2.0 if rng.random() < 0.5 else 1.0)
        mv[c] = np.kron(coarse, np.ones((bs, bs)))[:H, :W]
    if kind == "frac":
        mv += rng.normal(0, 0.7, mv.shape)
    if kind == "mixed":
        k = int(rng.integers(1, 6))
        mv[rng.integers(0, 2, k), rng.integers(0, H, k), rng.integers(0, W, k)] = rng.choice([0.5, -0.5, 2.5, -7.25, 1e9, np.nan], size=k)
    sub, P1, P2, diag, passes, adp = 1, 6, 32, 1, 2, int(rng.random() < 0.2)
    want = oracle(fp["I1"], fp["I2"], mv, rx, ry, agg, sub, P1, P2, diag, passes, adp)
    bestD, minC, mvSub = ctx.calc_pyd_cost_sgm(fp["I1"], fp["I2"], mv, rx, ry, agg, sub, P1, P2, diag, passes, adp)
    ok = np.array_equal(minC, want["minC"]) and np.array_equal(bestD, want["bestD"]) and np.array_equal(mvSub, want["mvSub"], equal_nan=True)
    print(f"case {it}: {W}x{H} r=({rx},{ry}) agg={agg} blocks={bs} prior={kind} adaptive={adp}: {'ok' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        print("  minC diffs:", int((minC != want["minC"]).sum()), "bestD diffs:", int((bestD != want["bestD"]).sum()))
        sys.exit(1)
print("all cases agree")
