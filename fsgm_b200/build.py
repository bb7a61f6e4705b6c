"""Builds fsgm_b200/libfsgm.so (the product) with nvcc for sm_100a.  No GPU is needed to build."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfsgm.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=hidden", "-shared",
    "--fmad=false",            # fp64 geometry must not be contracted; integer kernels do not care
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src) + ".o")
        cmd = [NVCC] + [f for f in FLAGS if f != "-shared"] + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {os.path.basename(src)}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC"], check=True)
    return LIB


PROJ = os.path.join(HERE, "proj")
PROJ_LIB = os.path.join(HERE, "libfsgm_proj.so")
PROJ_BIN = os.path.join(HERE, "sgmof")


def build_proj(force: bool = False) -> str:
    """C++ facade of the reference's proj/ library (EpiSGM / PydSGM, KITTI flow PNG, calib reader) + the sgmof command line.
    Plain host C++ on top of libfsgm.so's C ABI; zlib for PNG."""
    srcs = [os.path.join(PROJ, f) for f in ("png_io.cpp", "flow_io.cpp", "sgm_facade.cpp")]
    main = os.path.join(PROJ, "sgmof_main.cpp")
    deps = srcs + [main, os.path.join(HERE, "..", "include", "fsgm_proj.hpp"), LIB]
    if not force and os.path.exists(PROJ_LIB) and os.path.exists(PROJ_BIN) and \
            all(os.path.getmtime(p) <= min(os.path.getmtime(PROJ_LIB), os.path.getmtime(PROJ_BIN)) for p in deps):
        return PROJ_BIN
    cxx = os.environ.get("FSGM_CXX", "g++")        # the system g++ (shared libstdc++), not $CXX: a static libstdc++ inside a dlopen-ed library breaks iostreams
    common = ["-O2", "-std=c++17", "-Wall", "-fPIC"]
    link = ["-L" + HERE, "-lfsgm", "-lz", "-Wl,-rpath,$ORIGIN"]
    subprocess.run([cxx] + common + ["-shared", "-o", PROJ_LIB] + srcs + link, check=True)
    subprocess.run([cxx] + common + ["-o", PROJ_BIN, main, "-L" + HERE, "-lfsgm_proj"] + link, check=True)
    return PROJ_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_proj(force="--force" in sys.argv))
