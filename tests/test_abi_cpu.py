"""The product library loads without a GPU and exports every symbol include/*.h declares (no compute calls here)."""
import ctypes
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        names += re.findall(r"FSGM_API\s+[\w\s\*]+?\b(fsgm_\w+)\s*\(", open(h).read())
    return sorted(set(names))


def test_build_and_exports():
    from fsgm_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert lib.fsgm_abi_version() >= 1


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fsgm_b200 import api
    with pytest.raises(api.FsgmError):
        api.Context(0)


def test_product_never_touches_oracle():
    """Nothing under fsgm_b200/ may import, link or load anything from oracle/."""
    for path in glob.glob(os.path.join(ROOT, "fsgm_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
            txt = open(path, errors="ignore").read()
            assert "pyoracle" not in txt and "fsgm_oracle" not in txt and "libref_" not in txt, path
    out = subprocess.run(["ldd", os.path.join(ROOT, "fsgm_b200", "libfsgm.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "fsgm_b200", "libfsgm.so")], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_glibc_rand_matches_libc():
    """The library's own implementation of glibc's TYPE_3 generator against the process's libc (srand/rand)."""
    import numpy as np
    from fsgm_b200 import api
    libc = ctypes.CDLL(None)
    for seed in (1, 0, 42, 2 ** 31 + 5):
        libc.srand(ctypes.c_uint(seed))
        want = np.array([libc.rand() for _ in range(2000)], np.int32)
        assert np.array_equal(api.glibc_rand(seed, 2000), want), seed


def test_multi_gpu_plan_helpers_are_host_only():
    """fsgm_shard_range / fsgm_dirsplit_plan (csrc/dist.cu) need no GPU: block partition of a batch, and the direction-split plan
    (slabs even and covering the image, every direction owned exactly once, u8 exchange exactly when every rank's directions fit a
    byte together)."""
    from fsgm_b200 import api
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            seen = [i for r in range(world) for i in api.shard_range(n, r, world)]
            assert seen == list(range(n))
    for world in (1, 2, 3, 4, 8):
        for paths in (4, 8):
            infos = [api.dirsplit_plan(1242, 375, 256, paths, 6, 64, r, world) for r in range(world)]
            dirs = sorted(d for i in infos for d in i.dirs[:i.n_dirs])
            assert dirs == ([0, 1, 2, 3, 4, 5, 6, 7] if paths == 8 else [0, 1, 4, 5])
            assert all(i.slab_pixels % 2 == 0 and i.padded_pixels == i.slab_pixels * world for i in infos)
            assert sum(i.n_pixels for i in infos) == 1242 * 375
            assert [i.first_pixel for i in infos] == [min(1242 * 375, r * infos[0].slab_pixels) for r in range(world)]
            kmax = -(-paths // world)
            assert all(i.exchange_u8 == int(kmax * (24 + 64) <= 255) for i in infos)
    assert api.dirsplit_plan(64, 64, 40, 8, 6, 64, 0, 8).exchange_u8 == 0          # label count not a multiple of 16
    assert api.dirsplit_plan(64, 64, 64, 8, 200, 250, 0, 8).exchange_u8 == 0       # mod-256 parameter domain
    with pytest.raises(api.FsgmError):
        api.dirsplit_plan(64, 64, 64, 5, 6, 64, 0, 2)
