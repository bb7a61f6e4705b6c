"""torchrun worker for tests/test_dist_gpu.py: direction-split epipolar SGM over NCCL vs the single-GPU gateway."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fsgm_b200 import api, synth           # noqa: E402
from fsgm_b200 import dist as fd           # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    ctx.use_torch_stream()
    be = fd.GpuBackend(ctx)
    ok = True
    for (W, H, D, paths) in ((97, 61, 64, 8), (130, 75, 256, 8), (64, 48, 128, 4), (1242, 375, 256, 8)):
        p = synth.epipolar_pair(W, H, D, seed=3)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        bestD, minC = fd.epi_direction_split(be, p, D, p["vMax"], 6, 64, paths=paths)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        b1, m1, _, _ = ctx.calc_cost_sgm(p["I1"], p["I2"], D, p["vMax"], p["Pd0"], p["dirn"], p["O"], 6, 64, opts=api.epi_opts(paths=paths))
        same = np.array_equal(bestD, b1) and np.array_equal(minC, m1)
        ok &= same
        if dist.get_rank() == 0:
            print(f"dirsplit {W}x{H} D={D} paths={paths} world={dist.get_world_size()}: {'OK' if same else 'MISMATCH'} ({dt * 1e3:.1f} ms incl. H2D)", flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
