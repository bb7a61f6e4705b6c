"""Multi-GPU harness for the fSGM hot path.  One process per GPU.

The product path lives in the C ABI (csrc/dist.cu): `fsgm_dist_init` gives every context an NCCL communicator and
`fsgm_calc_cost_sgm_dirsplit_dev` runs the direction split with the collectives issued from C++; `nccl_init()` below only carries
the 128-byte rendezvous id from rank 0 to the others over torch.distributed, and `shard_range` / `split_directions` /
`slab_pixels` are views of the library's own plan (`fsgm_shard_range`, `fsgm_dirsplit_plan`).

`epi_direction_split` is the same data flow written against a small backend interface with torch.distributed collectives: it is
what the CPU test-suite runs over gloo (world_size 2 and 4) with the oracle as compute backend, slab by slab against the
single-process oracle, and it can drive the library's stage seams on GPUs (GpuBackend).

Two partitionings, nothing else:

* batch of independent pairs  -> `shard_range`: rank r owns a contiguous block of pairs; no data-path collective.
* one large pair, scan directions split across ranks -> `epi_direction_split`: every rank builds the cost volume
  (cheap, redundant), aggregates only its own directions into a u16 partial volume, the partial volumes are summed
  with ONE reduce-scatter over pixel slabs (the u16 pairs viewed as 32-bit words: per-voxel totals are at most
  8*255 < 65536, so no carry ever crosses a half-word and the integer sum is exact), each rank runs WTA/subpixel on
  its slab and the 8 B/pixel outputs are all-gathered.  The reference's read of the *next pixel's* label 0 for
  argmin == dMax-1 (calc_cost_sgm.cpp:293-296) crosses slab boundaries: the first voxel of every slab is all-gathered
  (one u16 per rank) and handed to the previous rank's WTA.

The compute steps go through a small backend object so the same logic runs on GPUs (GpuBackend: the C ABI via
fsgm_b200.api) and, in the CPU test-suite, over gloo with a checker backend supplied by the test.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import api

ALL_DIRS_8 = (0, 1, 2, 3, 4, 5, 6, 7)       # L1(+x) L3(+y) L2(+x+y) L4(-x+y), then reversed (include/fsgm.h)
ALL_DIRS_4 = (0, 1, 4, 5)


def shard_range(n: int, rank: int, world: int) -> range:
    """contiguous block partition of n independent units (pairs) over `world` ranks (fsgm_shard_range)"""
    return api.shard_range(n, rank, world)


def split_directions(paths: int, rank: int, world: int):
    """this rank's scan directions (fsgm_dirsplit_plan)"""
    info = api.dirsplit_plan(2, 2, 16, paths, 0, 0, rank, world)
    return tuple(info.dirs[:info.n_dirs])


def slab_pixels(N: int, world: int) -> int:
    """pixels per rank after the reduction: equal slabs, even so that slab*D u16 values pack into whole u32 words"""
    return int(api.dirsplit_plan(N, 1, 16, 8, 0, 0, 0, world).slab_pixels)


def nccl_init(ctx, group=None):
    """Give `ctx` its NCCL communicator (fsgm_dist_init): rank 0 creates the rendezvous id, torch.distributed carries the 128
    bytes to the other ranks (any backend), every rank joins."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [api.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.dist_init(box[0], rank, world)


class GpuBackend:
    """Compute steps of the direction-split path on one GPU, through the C ABI."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)
        # the collectives and the torch fills around the C-ABI kernels run on torch's current stream: the kernels
        # must be ordered on that same stream, not on the context's private one
        with torch.cuda.device(self.device):
            ctx.use_torch_stream()

    def upload(self, pair):
        """copy a pair's arrays to this GPU once (device-resident measurements reuse the result)"""
        dev = {k: torch.from_numpy(np.ascontiguousarray(pair[k][None])).to(self.device) for k in ("I1", "I2", "Pd0", "dirn", "O")}
        dev["_device"] = True
        return dev

    def cost_volume(self, pair, D, vMax):
        if not pair.get("_device"):
            pair = self.upload(pair)
        I1, I2, Pd0, dirn, O = (pair[k] for k in ("I1", "I2", "Pd0", "dirn", "O"))
        _, H, W = I1.shape
        cen1 = torch.empty((1, H, W), dtype=torch.int32, device=self.device)
        cen2 = torch.empty_like(cen1)
        self.ctx.census_dev(I1, cen1)
        self.ctx.census_dev(I2, cen2)
        Cv = torch.empty((1, H, W, D), dtype=torch.uint8, device=self.device)
        self.ctx.epi_cost_dev(cen1, cen2, D, vMax, Pd0, dirn, O, None, Cv)
        return Cv[0], I1[0]

    def partial(self, Cvol, I1, P1, P2, dirs, n_pad):
        H, W, D = Cvol.shape
        key = (n_pad, D)
        if getattr(self, "_partial_key", None) != key:       # reused across calls: the kernel rewrites every real voxel,
            self._partial = torch.zeros(n_pad * D, dtype=torch.int16, device=self.device)   # the padding stays zero
            self._partial_key = key
        self.ctx.epi_partial_dev(Cvol, I1, P1, P2, dirs, self._partial)
        return self._partial

    def partial_u8(self, Cvol, I1, P1, P2, dirs, n_pad):
        H, W, D = Cvol.shape
        key = ("u8", n_pad, D)
        if getattr(self, "_partial8_key", None) != key:
            self._partial8 = torch.zeros(n_pad * D, dtype=torch.uint8, device=self.device)
            self._partial8_key = key
        self.ctx.epi_partial_u8_dev(Cvol, I1, P1, P2, dirs, self._partial8)
        return self._partial8

    def wta_slabs(self, slabs, n_slabs, next0, D, O_slab, vMax):
        n = slabs.numel() // (D * n_slabs)
        bestD = torch.empty(n, dtype=torch.int32, device=self.device)
        minC = torch.empty_like(bestD)
        self.ctx.epi_wta_slabs_dev(slabs, n_slabs, next0, D, O_slab, vMax, bestD, minC)
        return bestD, minC

    def wta(self, Sp_slab, next0, D, O_slab, vMax):
        n = Sp_slab.numel() // D
        bestD = torch.empty(n, dtype=torch.int32, device=self.device)
        minC = torch.empty_like(bestD)
        self.ctx.epi_wta_sp_dev(Sp_slab, next0, D, O_slab, vMax, bestD, minC)
        return bestD, minC

    def to_device(self, a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device)


def epi_direction_split(backend, pair, D, vMax, P1, P2, paths=8, group=None):
    """calc_cost_sgm for ONE pair with the scan directions split over the ranks of `group`.
    Returns (bestD, minC) as uint32 numpy arrays [H][W], identical on every rank and identical to the single-GPU call."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    H, W = pair["I1"].shape[-2:]
    N = H * W
    plan = api.dirsplit_plan(W, H, D, paths, P1, P2, rank, world)          # the library's own plan (csrc/dist.cu)
    slab, n_pad = int(plan.slab_pixels), int(plan.padded_pixels)
    Cvol, I1 = backend.cost_volume(pair, D, vMax)
    my_dirs = tuple(plan.dirs[:plan.n_dirs])
    use_u8 = bool(plan.exchange_u8) and hasattr(backend, "partial_u8") and os.environ.get("FSGM_DIRSPLIT_U16") != "1"
    if use_u8:
        # every rank's directions fit a byte together: exchange u8 pixel slabs with ONE all-to-all (half the bytes of the
        # u16 reduce-scatter) and sum the received slabs inside the WTA kernel
        part8 = backend.partial_u8(Cvol, I1, P1, P2, my_dirs, n_pad)
        recv = torch.empty_like(part8)
        dist.all_to_all_single(recv, part8, group=group)                  # recv[j] = rank j's partial of OUR slab
        first = recv.view(world, slab * D)[:, 0].to(torch.int32).sum().to(torch.int32).reshape(1)
        firsts = [torch.empty(1, dtype=torch.int32, device=recv.device) for _ in range(world)]
        dist.all_gather(firsts, first, group=group)
        next0 = firsts[rank + 1].to(torch.int16) if rank + 1 < world else None
        O_slab = _o_slab(backend, pair, N, slab, n_pad, rank)
        bestD, minC = backend.wta_slabs(recv, world, next0, D, O_slab, vMax)
        return _gather_outputs(pair, bestD, minC, world, N, H, W, group)
    part = backend.partial(Cvol, I1, P1, P2, my_dirs, n_pad)    # int16 bits = u16
    words = part.view(torch.int32)                                                               # two u16 per word
    per = slab * D // 2
    if dist.get_backend(group) == "nccl":
        mine = torch.empty(per, dtype=torch.int32, device=words.device)
        dist.reduce_scatter_tensor(mine, words, op=dist.ReduceOp.SUM, group=group)
    else:                                                   # gloo has no reduce-scatter: all-reduce, keep our slab
        dist.all_reduce(words, op=dist.ReduceOp.SUM, group=group)
        mine = words[rank * per:(rank + 1) * per].clone()
    Sp_slab = mine.view(torch.int16)
    # one u16 per rank, exchanged as int32 (gloo has no 16-bit all_gather); totals are < 2^15 so the round trip is exact
    firsts = [torch.empty(1, dtype=torch.int32, device=Sp_slab.device) for _ in range(world)]
    dist.all_gather(firsts, Sp_slab[:1].to(torch.int32) & 0xFFFF, group=group)
    next0 = firsts[rank + 1].to(torch.int16) if rank + 1 < world else None
    O_slab = _o_slab(backend, pair, N, slab, n_pad, rank)
    bestD, minC = backend.wta(Sp_slab, next0, D, O_slab, vMax)
    return _gather_outputs(pair, bestD, minC, world, N, H, W, group)


def _o_slab(backend, pair, N, slab, n_pad, rank):
    if pair.get("_device"):
        O_slab = torch.zeros(slab, dtype=torch.float64, device=pair["O"].device)
        lo, hi = rank * slab, min(N, (rank + 1) * slab)
        if hi > lo:
            O_slab[:hi - lo] = pair["O"].reshape(-1)[lo:hi]
        return O_slab
    O_pad = np.zeros(n_pad, np.float64)
    O_pad[:N] = pair["O"].reshape(-1)
    return backend.to_device(O_pad[rank * slab:(rank + 1) * slab])


def _gather_outputs(pair, bestD, minC, world, N, H, W, group):
    out = torch.stack([bestD, minC])                                                            # [2][slab]
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out, group=group)
    full = torch.cat(gathered, dim=1)[:, :N]
    if pair.get("_device") and pair.get("_keep_on_device"):
        return full[0].reshape(H, W), full[1].reshape(H, W)          # int32 CUDA tensors (u32 bit patterns)
    full = full.cpu().numpy().view(np.uint32)
    return full[0].reshape(H, W), full[1].reshape(H, W)
