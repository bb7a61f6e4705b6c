"""torchrun worker for tests/test_dist_gpu.py: the direction-split path of the C ABI (fsgm_calc_cost_sgm_dirsplit_dev: NCCL issued
from C++) vs the single-GPU gateway, plus one case through the stage seams with torch.distributed collectives (GpuBackend)."""
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
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = api.Context(local)
    ctx.use_torch_stream()
    fd.nccl_init(ctx)
    ok = True
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[None])).cuda()
    # (W, H, D, paths, P1, P2, adaptive): u8 and u16 exchange, 4 paths (idle ranks when world > 4), mod-256 domain, adaptive P2,
    # a label count the fused cost kernel does not take, an image smaller than the rank count's slabs
    cases = ((97, 61, 64, 8, 6, 64, 0), (130, 75, 256, 8, 6, 32, 0), (64, 48, 128, 4, 6, 64, 0), (50, 33, 40, 8, 200, 250, 0),
             (71, 40, 64, 8, 6, 64, 1), (3, 1, 16, 8, 6, 64, 0), (1242, 375, 256, 8, 6, 64, 0))
    # every case twice: the peer-store form (sweeps write into the slab owners' memory over NVLink) and, with fsgm_tune key 4, the
    # NCCL exchange of partial volumes that stands behind it
    for no_p2p, (W, H, D, paths, P1, P2, adp) in [(m, cs) for m in (0, 1) for cs in cases]:
        ctx.tune(4, no_p2p)
        p = synth.epipolar_pair(W, H, D, seed=3)
        o = api.epi_opts(paths=paths, adaptive_p2=adp)
        dev = [t(p[k]) for k in ("I1", "I2", "Pd0", "dirn", "O")]
        bestD = torch.empty((1, H, W), dtype=torch.int32, device="cuda"); minC = torch.empty_like(bestD)
        b1 = torch.empty_like(bestD); m1 = torch.empty_like(bestD)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        ctx.calc_cost_sgm_dirsplit_dev(dev[0], dev[1], D, p["vMax"], dev[2], dev[3], dev[4], P1, P2, bestD, minC, opts=o)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ctx.calc_cost_sgm_dev(dev[0], dev[1], D, p["vMax"], dev[2], dev[3], dev[4], P1, P2, b1, m1, opts=o)
        same = bool(torch.equal(bestD, b1) and torch.equal(minC, m1))
        ok &= same
        plan = api.dirsplit_plan(W, H, D, paths, P1, P2, rank, world)
        if rank == 0:
            form = ("peer stores" if (not no_p2p and not adp and D % 16 == 0 and P1 + P2 + 24 <= 255) else
                    "u8 exchange" if plan.exchange_u8 else "u16 reduce-scatter")
            print(f"dirsplit(C ABI) {W}x{H} D={D} paths={paths} P=({P1},{P2}) adaptive={adp} world={world} {form}: "
                  f"{'OK' if same else 'MISMATCH'} ({dt * 1e3:.1f} ms)", flush=True)
    ctx.tune(4, 0)
    # the stage seams with torch.distributed collectives
    be = fd.GpuBackend(ctx)
    W, H, D, paths = 130, 75, 256, 8
    p = synth.epipolar_pair(W, H, D, seed=3)
    bestD, minC = fd.epi_direction_split(be, p, D, p["vMax"], 6, 64, paths=paths)
    b1, m1, _, _ = ctx.calc_cost_sgm(p["I1"], p["I2"], D, p["vMax"], p["Pd0"], p["dirn"], p["O"], 6, 64, opts=api.epi_opts(paths=paths))
    same = np.array_equal(bestD, b1) and np.array_equal(minC, m1)
    ok &= same
    if rank == 0:
        print(f"dirsplit(stage seams + torch.distributed) {W}x{H} D={D}: {'OK' if same else 'MISMATCH'}", flush=True)
    # a context without a communicator refuses loudly
    c2 = api.Context(local)
    try:
        c2._l.fsgm_dist_adopt_comm  # symbol present
        dev = [t(p[k]) for k in ("I1", "I2", "Pd0", "dirn", "O")]
        bb = torch.empty((1, H, W), dtype=torch.int32, device="cuda")
        c2.dist_allgather_u32(bb, bb)
        ok = False
        print("MISMATCH: allgather without a communicator did not fail", flush=True)
    except api.FsgmError as e:
        assert e.code == api.FSGM_ERR_NCCL
    c2.close()
    tt = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MIN)
    dist.barrier()
    ctx.dist_finalize()
    dist.destroy_process_group()
    sys.exit(0 if int(tt.item()) == 1 else 1)


if __name__ == "__main__":
    main()
