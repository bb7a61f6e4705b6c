#!/usr/bin/env python
"""Secondary measurement (BASELINE.json configs[4], second half): ONE 3840x2160 pair, 256 labels, 8 paths, the scan
directions split across the GPUs of one box and the per-direction volumes summed by an NCCL reduce-scatter.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 bench_dirsplit.py

Inputs are device-resident; timing = CUDA events, max over ranks.  Prints one JSON line on rank 0.  With N = 1 the same
script times the plain single-GPU call (fsgm_calc_cost_sgm_dev) for comparison.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fsgm_b200 import api, synth           # noqa: E402
from fsgm_b200 import dist as fd           # noqa: E402

W, H, D, P1, P2 = 3840, 2160, 256, 6, 64


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    ctx.use_torch_stream()
    p = synth.epipolar_pair(W, H, D, seed=9)
    steps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world == 1:
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a[None])).cuda()
        I1, I2, Pd0, dirn, O = (t(p[k]) for k in ("I1", "I2", "Pd0", "dirn", "O"))
        b = torch.empty((1, H, W), dtype=torch.int32, device="cuda"); m = torch.empty_like(b)
        f = lambda: ctx.calc_cost_sgm_dev(I1, I2, D, 0.3, Pd0, dirn, O, P1, P2, b, m, opts=api.epi_opts(paths=8))
    else:
        be = fd.GpuBackend(ctx)
        dev = be.upload(p)
        dev["_keep_on_device"] = True
        f = lambda: fd.epi_direction_split(be, dev, D, 0.3, P1, P2, paths=8)
    f(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(steps):
        f()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if int(os.environ.get("RANK", "0")) == 0:
        N = W * H
        print(json.dumps({"metric": "ms per 3840x2160 pair (256 labels, 8 paths), directions split over GPUs",
                          "n_gpus": world, "value": float(ms.item()), "unit": "ms", "higher_is_better": False,
                          "gde_per_s": N * D / (float(ms.item()) * 1e-3) / 1e9,
                          "exchange": ("none" if world == 1 else
                                       "u16 reduce-scatter" if (world < 4 or os.environ.get("FSGM_DIRSPLIT_U16") == "1") else "u8 all-to-all"),
                          "nvlink_bytes_per_rank": 0 if world == 1 else int((world - 1) / world * N * D *
                                                   (2 if (world < 4 or os.environ.get("FSGM_DIRSPLIT_U16") == "1") else 1))}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
