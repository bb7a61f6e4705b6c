"""Host-side multi-GPU logic over gloo, world_size 2, on CPU: partitioning helpers and the direction-split data flow
(packing of u16 pairs into 32-bit words, slab reduce, cross-slab label-0 exchange, output gather).  The compute steps
are played by the CPU oracle (tests may use it); the result must equal the single-process oracle bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fsgm_b200 import dist as fd
from fsgm_b200 import synth


def test_shard_range_partitions():
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            seen = [i for r in range(world) for i in fd.shard_range(n, r, world)]
            assert seen == list(range(n))
            sizes = [len(fd.shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_split_directions_cover():
    for paths, want in ((8, 8), (4, 4)):
        for world in (1, 2, 3, 4, 8):
            got = sorted(d for r in range(world) for d in fd.split_directions(paths, r, world))
            assert len(got) == want and len(set(got)) == want
    assert fd.slab_pixels(11, 2) % 2 == 0 and fd.slab_pixels(11, 2) * 2 >= 11


class OracleBackend:
    """same interface as fsgm_b200.dist.GpuBackend, computed by oracle/fsgm_oracle.c on the CPU"""

    def __init__(self, po):
        self.po = po

    def cost_volume(self, pair, D, vMax):
        r = self.po.port_epi(pair["I1"], pair["I2"], D, vMax, pair["Pd0"], pair["dirn"], pair["O"], 6, 32, paths=4)
        return r["C"], pair["I1"]

    def partial(self, Cvol, I1, P1, P2, dirs, n_pad):
        H, W, D = Cvol.shape
        acc = np.zeros(n_pad * D, np.uint16)
        for r in dirs:
            acc[:H * W * D] += self.po.port_sweep1d(Cvol, I1, P1, P2, r).reshape(-1)
        return torch.from_numpy(acc.view(np.int16))

    def partial_u8(self, Cvol, I1, P1, P2, dirs, n_pad):
        H, W, D = Cvol.shape
        acc = np.zeros(n_pad * D, np.uint8)
        for r in dirs:
            acc[:H * W * D] += self.po.port_sweep1d(Cvol, I1, P1, P2, r).reshape(-1)
        return torch.from_numpy(acc)

    def wta_slabs(self, slabs, n_slabs, next0, D, O_slab, vMax):
        sp = slabs.numpy().reshape(n_slabs, -1).astype(np.uint16).sum(0).astype(np.uint16)
        return self.wta(torch.from_numpy(sp.view(np.int16)), next0, D, O_slab, vMax)

    def wta(self, Sp_slab, next0, D, O_slab, vMax):
        sp = Sp_slab.numpy().view(np.uint16).astype(np.uint32).reshape(1, -1, D)
        n = sp.shape[1]
        ext = np.zeros((1, n + 1, D), np.uint32)               # the oracle reads pixel n's label 0 for the last pixel
        ext[0, :n] = sp[0]
        if next0 is not None:
            ext[0, n, 0] = int(next0.numpy().view(np.uint16)[0])
        bestD, minC = self.po.port_epi_wta(np.ascontiguousarray(ext))
        bestD, minC = np.ascontiguousarray(bestD[0, :n]), np.ascontiguousarray(minC[0, :n])
        O = np.ascontiguousarray(O_slab.numpy())
        self.po._port().orc_vz_to_disp(bestD.ctypes.data_as(C.c_void_p), n, 1, O.ctypes.data_as(C.c_void_p), C.c_double(vMax), D)
        return torch.from_numpy(bestD.view(np.int32)), torch.from_numpy(minC.view(np.int32))

    def to_device(self, a):
        return torch.from_numpy(np.ascontiguousarray(a))


def _worker(rank, world, port, W, H, D, paths, P2, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    p = synth.epipolar_pair(W, H, D, seed=5)
    bestD, minC = fd.epi_direction_split(OracleBackend(po), p, D, p["vMax"], 6, P2, paths=paths)
    q.put((rank, bestD, minC))
    dist.barrier()
    dist.destroy_process_group()


# world 2, P2 = 64: four directions per rank do not fit a byte -> u16 reduce path; the other cases take the u8 all-to-all path
@pytest.mark.parametrize("W,H,D,paths,P2,world", [(37, 23, 16, 8, 64, 2), (20, 15, 8, 4, 64, 2), (29, 19, 16, 8, 32, 2), (31, 17, 16, 8, 64, 4)])
def test_direction_split_gloo(oracle, W, H, D, paths, P2, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, D, paths, P2, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = synth.epipolar_pair(W, H, D, seed=5)
    want = oracle.port_epi(p["I1"], p["I2"], D, p["vMax"], p["Pd0"], p["dirn"], p["O"], 6, P2, paths=paths)
    for rank, bestD, minC in res:
        assert np.array_equal(minC, want["minC"]), rank
        a, b = bestD.copy(), want["bestD"].copy()
        if want["Sp"][-1, -1].argmin() == D - 1:
            a[-1, -1] = b[-1, -1] = 0
        assert np.array_equal(a, b), rank
