"""ctypes binding of libfsgm.so (include/fsgm.h) — the harness tests/ and bench.py drive the product through.

The functions keep the reference's gateway names and argument order (calc_cost_sgm.cpp:539 etc.), with numpy
host arrays in the reference's layout: row-major, x fastest, images uint8 [H][W], two-plane doubles [2][H][W].
The *_dev methods take torch CUDA tensors (device memory plumbing only) and pass raw device pointers through.

There is no fallback: importing this module without a built libfsgm.so raises, and every call raises
FsgmError on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfsgm.so")

FSGM_OK, FSGM_ERR_ARG, FSGM_ERR_DOMAIN, FSGM_ERR_CUDA, FSGM_ERR_NOMEM, FSGM_ERR_NCCL = 0, -1, -2, -3, -4, -5


class FsgmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fsgm error {code}: {msg}")
        self.code = code


class EpiOpts(C.Structure):
    _fields_ = [("paths", C.c_int), ("total_pass", C.c_int), ("subpixel", C.c_int),
                ("adaptive_p2", C.c_int), ("vz_to_disp", C.c_int), ("fb_check", C.c_int), ("fb_thr", C.c_int)]


def load_library() -> C.CDLL:
    global _LIB_PATH
    _LIB_PATH = os.environ.get("FSGM_LIB", _LIB_PATH)          # A/B builds of the same ABI (kernel experiments)
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: build it with `python fsgm_b200/build.py` "
                          "(there is no CPU fallback for the fSGM hot path)")
    lib = C.CDLL(_LIB_PATH)
    lib.fsgm_last_error.restype = C.c_char_p
    lib.fsgm_launch_count.restype = C.c_uint64
    lib.fsgm_scratch_bytes.restype = C.c_size_t
    lib.fsgm_stage_name.restype = C.c_char_p
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def _hp(a: np.ndarray, dtype, shape=None):
    """host pointer of a C-contiguous numpy array of the right dtype"""
    if a is None:
        return None
    if a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise TypeError(f"expected C-contiguous {np.dtype(dtype)} array, got {a.dtype}")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a.ctypes.data_as(C.c_void_p)


def _dp(t):
    """device pointer of a contiguous torch CUDA tensor"""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise TypeError("expected a contiguous CUDA tensor")
    return C.c_void_p(t.data_ptr())


class PydOpts(C.Structure):
    """constants of the MATLAB pyramid driver (pyramidal_sgm.m:12-22)"""
    _fields_ = [("numPyd", C.c_int), ("P1", C.c_int), ("P2", C.c_int), ("aggHalfWinSize", C.c_int),
                ("verSearchHalfWinSize", C.c_int), ("horSearchHalfWinSize", C.c_int), ("enableDiagonal", C.c_int),
                ("totalPass", C.c_int), ("adaptiveP2", C.c_int)]


def pyd_opts(numPyd=5, P1=6, P2=32, agg=2, ver=5, hor=5, diag=1, passes=2, adaptive=0) -> PydOpts:
    return PydOpts(numPyd, P1, P2, agg, ver, hor, diag, passes, adaptive)


def pyramid_dims(W: int, H: int, numPyd: int):
    ws, hs = (C.c_int * numPyd)(), (C.c_int * numPyd)()
    rc = lib().fsgm_pyramid_dims(int(W), int(H), int(numPyd), ws, hs)
    if rc:
        raise FsgmError(rc, "fsgm_pyramid_dims")
    return list(ws), list(hs)


class DirsplitInfo(C.Structure):
    """fsgm_dirsplit_info: the plan of the direction-split path for one rank (include/fsgm.h)"""
    _fields_ = [("slab_pixels", C.c_size_t), ("padded_pixels", C.c_size_t), ("first_pixel", C.c_size_t), ("n_pixels", C.c_size_t),
                ("dirs", C.c_int * 8), ("n_dirs", C.c_int), ("exchange_u8", C.c_int)]


def shard_range(n: int, rank: int, world: int) -> range:
    """fsgm_shard_range: contiguous block of n independent units (pairs) owned by `rank`"""
    lo, cnt = C.c_int(), C.c_int()
    rc = lib().fsgm_shard_range(int(n), int(rank), int(world), C.byref(lo), C.byref(cnt))
    if rc != FSGM_OK:
        raise FsgmError(rc, "fsgm_shard_range")
    return range(lo.value, lo.value + cnt.value)


def dirsplit_plan(W: int, H: int, D: int, paths: int, P1: int, P2: int, rank: int, world: int) -> DirsplitInfo:
    info = DirsplitInfo()
    rc = lib().fsgm_dirsplit_plan(int(W), int(H), int(D), int(paths), int(P1), int(P2), int(rank), int(world), C.byref(info))
    if rc != FSGM_OK:
        raise FsgmError(rc, "fsgm_dirsplit_plan")
    return info


DIST_ID_BYTES = 128


def dist_unique_id() -> bytes:
    """fsgm_dist_unique_id: the NCCL rendezvous id rank 0 creates and hands to the other ranks"""
    buf = C.create_string_buffer(DIST_ID_BYTES)
    rc = lib().fsgm_dist_unique_id(buf, C.c_size_t(DIST_ID_BYTES))
    if rc != FSGM_OK:
        raise FsgmError(rc, "fsgm_dist_unique_id (is libnccl.so.2 loadable?)")
    return buf.raw


class NgOpts(C.Structure):
    _fields_ = [("seed", C.c_uint), ("rand_stream", C.c_void_p)]


def glibc_rand(seed: int, count: int) -> np.ndarray:
    """srand(seed); [rand() for _ in range(count)] — the library's own implementation of glibc's generator"""
    out = np.empty(count, np.int32)
    rc = lib().fsgm_glibc_rand_fill(C.c_uint(seed), C.c_size_t(count), out.ctypes.data_as(C.c_void_p))
    if rc != FSGM_OK:
        raise FsgmError(rc, "fsgm_glibc_rand_fill")
    return out


def epi_opts(paths=4, total_pass=2, subpixel=1, adaptive_p2=0, vz_to_disp=1, fb_check=0, fb_thr=2) -> EpiOpts:
    return EpiOpts(paths, total_pass, subpixel, adaptive_p2, vz_to_disp, fb_check, fb_thr)


class Context:
    """One per GPU (fsgm_create / fsgm_destroy)."""

    def __init__(self, device: int = 0):
        self._l = lib()
        self._h = C.c_void_p()
        rc = self._l.fsgm_create(int(device), C.byref(self._h))
        if rc != FSGM_OK:
            raise FsgmError(rc, f"fsgm_create(device={device}) failed — an sm_100 GPU is required, there is no CPU path")
        self.device = device

    def close(self):
        if self._h:
            self._l.fsgm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != FSGM_OK:
            raise FsgmError(rc, self._l.fsgm_last_error(self._h).decode())

    def set_stream(self, cuda_stream_ptr: int | None):
        self._ck(self._l.fsgm_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def use_torch_stream(self):
        import torch
        ptr = torch.cuda.current_stream(self.device).cuda_stream
        # torch's default stream is the NULL handle, which fsgm_set_stream reads as "back to the context's own
        # stream"; cudaStreamLegacy (0x1) names the same default stream explicitly.
        self.set_stream(ptr if ptr else 1)

    def tune(self, key: int, value: int):
        self._ck(self._l.fsgm_tune(self._h, int(key), int(value)))

    # ------------------------------------------------------------------ multi-GPU (NCCL communicator owned by the context)
    def dist_init(self, unique_id: bytes, rank: int, world: int):
        """fsgm_dist_init: collective over all ranks (ncclCommInitRank on this context's device)"""
        self._ck(self._l.fsgm_dist_init(self._h, C.c_char_p(unique_id), C.c_size_t(len(unique_id)), int(rank), int(world)))

    def dist_finalize(self):
        self._ck(self._l.fsgm_dist_finalize(self._h))

    def dist_allgather_u32(self, send, recv):
        self._ck(self._l.fsgm_dist_allgather_u32(self._h, _dp(send), C.c_size_t(send.numel()), _dp(recv)))

    def calc_cost_sgm_dirsplit_dev(self, I1, I2, dMax, vMax, Pd0, dirn, O, P1, P2, bestD, minC, opts=None):
        """ONE pair ([1][H][W] tensors, the same on every rank), scan directions split over the ranks of the communicator;
        bestD / minC are complete and identical on every rank"""
        H, W = I1.shape[-2:]
        self._ck(self._l.fsgm_calc_cost_sgm_dirsplit_dev(
            self._h, _dp(I1), _dp(I2), W, H, int(dMax), C.c_double(vMax), _dp(Pd0), _dp(dirn), _dp(O), int(P1), int(P2),
            C.byref(opts) if opts is not None else None, _dp(bestD), _dp(minC)))

    def epi_wave_pairs(self, W, D, P1, P2, opts=None) -> int:
        """pairs per wave of the cluster kernels for this shape (0 = generic kernels only)"""
        return int(self._l.fsgm_epi_wave_pairs(self._h, int(W), int(D), int(P1), int(P2), C.byref(opts) if opts is not None else None))

    def synchronize(self):
        self._ck(self._l.fsgm_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._l.fsgm_launch_count(self._h))

    @property
    def scratch_bytes(self) -> int:
        return int(self._l.fsgm_scratch_bytes(self._h))

    # ------------------------------------------------------------------ per-stage device timing
    def profile(self, on: bool):
        self._ck(self._l.fsgm_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._ck(self._l.fsgm_profile_reset(self._h))

    def profile_read(self) -> dict:
        """{stage name: (milliseconds, kernel launches)} accumulated since the last reset"""
        out = {}
        for st in range(self._l.fsgm_stage_count()):
            ms, n = C.c_double(), C.c_uint64()
            self._ck(self._l.fsgm_profile_read(self._h, st, C.byref(ms), C.byref(n)))
            if n.value:
                out[self._l.fsgm_stage_name(st).decode()] = (ms.value, n.value)
        return out

    # ------------------------------------------------------------------ gateway 1 (host arrays)
    def calc_cost_sgm(self, I1, I2, dMax, vMax, pixelPosD0, normlizeDirection, offsetFromPosD0, P1, P2, opts=None):
        """[bestD, minC, conf, bestD2] = calc_cost_sgm(...)  (calc_cost_sgm.cpp:17, :539-598)"""
        H, W = I1.shape
        bestD, minC = np.empty((H, W), np.uint32), np.empty((H, W), np.uint32)
        conf, bestD2 = np.empty((H, W), np.uint8), np.empty((H, W), np.uint32)
        self._ck(self._l.fsgm_calc_cost_sgm(
            self._h, _hp(I1, np.uint8), _hp(I2, np.uint8, (H, W)), W, H, int(dMax), C.c_double(vMax),
            _hp(pixelPosD0, np.float64, (2, H, W)), _hp(normlizeDirection, np.float64, (2, H, W)),
            _hp(offsetFromPosD0, np.float64, (H, W)), int(P1), int(P2), C.byref(opts) if opts is not None else None,
            _hp(bestD, np.uint32), _hp(minC, np.uint32), _hp(conf, np.uint8), _hp(bestD2, np.uint32)))
        return bestD, minC, conf, bestD2

    def calc_cost_sgm_batch(self, I1, I2, dMax, vMax, pixelPosD0, normlizeDirection, offsetFromPosD0, P1, P2, opts=None,
                            out=None, asynchronous=False):
        """Batch form; with asynchronous=True the call only enqueues (keep the arrays alive and call synchronize())."""
        n, H, W = I1.shape
        bestD, minC = out if out is not None else (np.empty((n, H, W), np.uint32), np.empty((n, H, W), np.uint32))
        fn = self._l.fsgm_calc_cost_sgm_batch_async if asynchronous else self._l.fsgm_calc_cost_sgm_batch
        self._ck(fn(
            self._h, n, _hp(I1, np.uint8), _hp(I2, np.uint8, (n, H, W)), W, H, int(dMax), C.c_double(vMax),
            _hp(pixelPosD0, np.float64, (n, 2, H, W)), _hp(normlizeDirection, np.float64, (n, 2, H, W)),
            _hp(offsetFromPosD0, np.float64, (n, H, W)), int(P1), int(P2), C.byref(opts) if opts is not None else None,
            _hp(bestD, np.uint32), _hp(minC, np.uint32)))
        return bestD, minC

    # ------------------------------------------------------------------ device-resident forms (torch tensors)
    def calc_cost_sgm_dev(self, I1, I2, dMax, vMax, Pd0, dirn, O, P1, P2, bestD, minC, opts=None):
        n, H, W = I1.shape
        self._ck(self._l.fsgm_calc_cost_sgm_dev(
            self._h, n, _dp(I1), _dp(I2), W, H, int(dMax), C.c_double(vMax), _dp(Pd0), _dp(dirn), _dp(O), int(P1), int(P2),
            C.byref(opts) if opts is not None else None, _dp(bestD), _dp(minC)))

    def forward_backward_check_dev(self, bestD, Pd0, dirn, O, vMax, n, conf, bestD2, thr=2, use_vzind=1):
        """forward_backward_check (calc_cost_sgm.cpp:488-536) on the x256 label map (before the vz conversion)"""
        npairs, H, W = bestD.shape
        self._ck(self._l.fsgm_forward_backward_check_dev(self._h, npairs, _dp(bestD), W, H, _dp(Pd0), _dp(dirn), _dp(O),
                                                         C.c_double(vMax), int(n), int(thr), int(use_vzind), _dp(conf), _dp(bestD2)))

    def convert_vzind_to_disp_dev(self, bestD, O, vMax, n):
        """convert_vzInd_to_disp (calc_cost_sgm.cpp:414-426), in place"""
        npairs, H, W = bestD.shape
        self._ck(self._l.fsgm_convert_vzind_to_disp_dev(self._h, npairs, _dp(bestD), W, H, _dp(O), C.c_double(vMax), int(n)))

    def census_dev(self, img, cen):
        n, H, W = img.shape
        self._ck(self._l.fsgm_census_dev(self._h, n, _dp(img), W, H, _dp(cen)))

    def epi_cost_dev(self, cen1, cen2, dMax, vMax, Pd0, dirn, O, raw, Cvol):
        n, H, W = cen1.shape
        self._ck(self._l.fsgm_epi_cost_dev(self._h, n, _dp(cen1), _dp(cen2), W, H, int(dMax), C.c_double(vMax),
                                           _dp(Pd0), _dp(dirn), _dp(O), _dp(raw), _dp(Cvol)))

    def sweep_dev(self, Cvol, I1, P1, P2, direction, L, adaptive_thr=0):
        n, H, W, D = Cvol.shape
        self._ck(self._l.fsgm_sweep_dev(self._h, n, _dp(Cvol), _dp(I1), W, H, D, int(P1), int(P2), int(adaptive_thr),
                                        int(direction), _dp(L)))

    def epi_aggregate_dev(self, Cvol, I1, P1, P2, O, vMax, bestD, minC, Sp=None, opts=None):
        n, H, W, D = Cvol.shape
        self._ck(self._l.fsgm_epi_aggregate_dev(self._h, n, _dp(Cvol), _dp(I1), W, H, D, int(P1), int(P2),
                                                C.byref(opts) if opts is not None else None, _dp(Sp), _dp(O),
                                                C.c_double(vMax), _dp(bestD), _dp(minC)))

    def epi_partial_dev(self, Cvol, I1, P1, P2, directions, Sp_partial, adaptive_p2=0):
        """sweeps for `directions` summed into a u16 partial volume (Cvol: [H][W][D] u8, single pair)"""
        H, W, D = Cvol.shape
        dirs = (C.c_int * max(1, len(directions)))(*[int(d) for d in directions])
        self._ck(self._l.fsgm_epi_partial_dev(self._h, _dp(Cvol), _dp(I1), W, H, D, int(P1), int(P2), int(adaptive_p2),
                                              dirs, len(directions), _dp(Sp_partial)))

    def epi_partial_u8_dev(self, Cvol, I1, P1, P2, directions, partial, adaptive_p2=0):
        H, W, D = Cvol.shape
        dirs = (C.c_int * max(1, len(directions)))(*[int(d) for d in directions])
        self._ck(self._l.fsgm_epi_partial_u8_dev(self._h, _dp(Cvol), _dp(I1), W, H, D, int(P1), int(P2), int(adaptive_p2),
                                                 dirs, len(directions), _dp(partial)))

    def epi_wta_slabs_dev(self, slabs, n_slabs, next0, D, O, vMax, bestD, minC, subpixel=1, vz_to_disp=1):
        n_pixels = slabs.numel() // (D * n_slabs)
        self._ck(self._l.fsgm_epi_wta_slabs_dev(self._h, _dp(slabs), int(n_slabs), _dp(next0), C.c_size_t(n_pixels), int(D),
                                                int(subpixel), int(vz_to_disp), _dp(O), C.c_double(vMax), _dp(bestD), _dp(minC)))

    def epi_wta_sp_dev(self, Sp, next0, D, O, vMax, bestD, minC, subpixel=1, vz_to_disp=1):
        n_pixels = Sp.numel() // D
        self._ck(self._l.fsgm_epi_wta_sp_dev(self._h, _dp(Sp), _dp(next0), C.c_size_t(n_pixels), int(D), int(subpixel),
                                             int(vz_to_disp), _dp(O), C.c_double(vMax), _dp(bestD), _dp(minC)))

    # ------------------------------------------------------------------ gateway 2: calc_pyd_cost_sgm
    def calc_pyd_cost_sgm(self, I1, I2, preMv, halfSearchWinSizeX, halfSearchWinSizeY, aggHalfWinSize, subPixelRefine,
                          P1, P2, enableDiagnalPath=1, totalPass=2, adpativeP2=0):
        """[bestD, minC, mvSub] = calc_pyd_cost_sgm(...)  (calc_pyd_cost_sgm.cpp:16, :439-510)"""
        H, W = I1.shape
        _, mvH, mvW = preMv.shape
        bestD, minC = np.empty((H, W), np.uint32), np.empty((H, W), np.uint32)
        mvSub = np.empty((2, H, W), np.float64)
        self._ck(self._l.fsgm_calc_pyd_cost_sgm(
            self._h, _hp(I1, np.uint8), _hp(I2, np.uint8, (H, W)), W, H, _hp(preMv, np.float64), mvW, mvH,
            int(halfSearchWinSizeX), int(halfSearchWinSizeY), int(aggHalfWinSize), int(subPixelRefine), int(P1), int(P2),
            int(enableDiagnalPath), int(totalPass), int(adpativeP2), _hp(bestD, np.uint32), _hp(minC, np.uint32),
            _hp(mvSub, np.float64)))
        return bestD, minC, mvSub

    def calc_pyd_cost_sgm_dev(self, I1, I2, preMv, rx, ry, agg, sub, P1, P2, diag, passes, adaptive, bestD, minC, mvSub):
        n, H, W = I1.shape
        _, _, mvH, mvW = preMv.shape
        self._ck(self._l.fsgm_calc_pyd_cost_sgm_dev(self._h, n, _dp(I1), _dp(I2), W, H, _dp(preMv), mvW, mvH, int(rx), int(ry),
                                                    int(agg), int(sub), int(P1), int(P2), int(diag), int(passes), int(adaptive),
                                                    _dp(bestD), _dp(minC), _dp(mvSub)))

    def pyd_cost_dev(self, cen1, cen2, preMv, agg, rx, ry, Cvol):
        n, H, W = cen1.shape
        _, _, mvH, mvW = preMv.shape
        self._ck(self._l.fsgm_pyd_cost_dev(self._h, n, _dp(cen1), _dp(cen2), W, H, _dp(preMv), mvW, mvH, int(agg), int(rx), int(ry),
                                           _dp(Cvol)))

    def pyd_sweep_dev(self, Cvol, I1, preMv, rx, ry, P1, P2, adaptive, direction, L):
        n, H, W, D = Cvol.shape
        _, _, mvH, mvW = preMv.shape
        self._ck(self._l.fsgm_pyd_sweep_dev(self._h, n, _dp(Cvol), _dp(I1), _dp(preMv), mvW, mvH, W, H, int(rx), int(ry),
                                            int(P1), int(P2), int(adaptive), int(direction), _dp(L)))

    def pyd_aggregate_dev(self, Cvol, I1, preMv, rx, ry, sub, P1, P2, diag, passes, adaptive, bestD, minC, mvSub, Sp=None):
        n, H, W, D = Cvol.shape
        _, _, mvH, mvW = preMv.shape
        self._ck(self._l.fsgm_pyd_aggregate_dev(self._h, n, _dp(Cvol), _dp(I1), _dp(preMv), mvW, mvH, W, H, int(rx), int(ry),
                                                int(sub), int(P1), int(P2), int(diag), int(passes), int(adaptive), _dp(Sp),
                                                _dp(bestD), _dp(minC), _dp(mvSub)))

    # ------------------------------------------------------------------ dense epipolar prologue / epilogue
    @staticmethod
    def _geo_args(F, Hm, epi, direction, n):
        F = np.ascontiguousarray(np.asarray(F, np.float64).reshape(n, 9))
        Hm = np.ascontiguousarray(np.asarray(Hm, np.float64).reshape(n, 9))
        epi = np.ascontiguousarray(np.asarray(epi, np.float64).reshape(n, 2))
        dr = np.ascontiguousarray(np.asarray(direction, np.int32).reshape(n))
        return F, Hm, epi, dr

    def epipolar_geometry_dev(self, F, Hm, epi, direction, Pd0, dirn, O, Rflow=None):
        """epipolar_geometry.m:104-119 + rotation_motion.m on the device; F, Hm, epi, direction are host arrays"""
        n, H, W = O.shape
        F, Hm, epi, dr = self._geo_args(F, Hm, epi, direction, n)
        self._ck(self._l.fsgm_epipolar_geometry_dev(self._h, n, _hp(F, np.float64), _hp(Hm, np.float64), _hp(epi, np.float64),
                                                    _hp(dr, np.int32), W, H, _dp(Pd0), _dp(dirn), _dp(O), _dp(Rflow)))

    def epipolar_flow_dev(self, bestD, dirn, Rflow, flow):
        n, H, W = bestD.shape
        self._ck(self._l.fsgm_epipolar_flow_dev(self._h, n, _dp(bestD), _dp(dirn), _dp(Rflow), W, H, _dp(flow)))

    def epipolar_sgm_of_dev(self, I0, I1, F, Hm, epi, direction, dMax, vMax, P1, P2, work, flow, minC, opts=None):
        n, H, W = I0.shape
        F, Hm, epi, dr = self._geo_args(F, Hm, epi, direction, n)
        self._l.fsgm_epipolar_sgm_of_work_bytes.restype = C.c_size_t
        if work.numel() * work.element_size() < self._l.fsgm_epipolar_sgm_of_work_bytes(n, W, H):
            raise ValueError("work buffer too small")
        self._ck(self._l.fsgm_epipolar_sgm_of_dev(self._h, n, _dp(I0), _dp(I1), W, H, _hp(F, np.float64), _hp(Hm, np.float64),
                                                  _hp(epi, np.float64), _hp(dr, np.int32), int(dMax), C.c_double(vMax), int(P1), int(P2),
                                                  C.byref(opts) if opts is not None else None, _dp(work), _dp(flow), _dp(minC)))

    def epipolar_sgm_of_batch(self, I0, I1, F, Hm, epi, direction, dMax, vMax, P1, P2, opts=None, out=None, asynchronous=False,
                              f32=False):
        """[flow, minC] = epipolar_sgm_of(...) from F, H, epipole on host images (epipolar_sgm_of.m:23-51 after the geometry fit).
        f32=True: flow as float32 [n][H][W][2] (CV_32FC2, proj/include/epi_sgm.h:6-12); minC may then be None in `out`."""
        n, H, W = I0.shape
        F, Hm, epi, dr = self._geo_args(F, Hm, epi, direction, n)
        if f32:
            flow, minC = out if out is not None else (np.empty((n, H, W, 2), np.float32), np.empty((n, H, W), np.uint32))
            fn, fdt = self._l.fsgm_epipolar_sgm_of_f32_batch_async, np.float32
        else:
            flow, minC = out if out is not None else (np.empty((n, 2, H, W), np.float64), np.empty((n, H, W), np.uint32))
            fn, fdt = self._l.fsgm_epipolar_sgm_of_batch_async, np.float64
        self._keep = (F, Hm, epi, dr)                    # host matrices are consumed at enqueue time, images are not
        self._ck(fn(
            self._h, n, _hp(I0, np.uint8), _hp(I1, np.uint8, (n, H, W)), W, H, _hp(F, np.float64), _hp(Hm, np.float64),
            _hp(epi, np.float64), _hp(dr, np.int32), int(dMax), C.c_double(vMax), int(P1), int(P2),
            C.byref(opts) if opts is not None else None, _hp(flow, fdt), _hp(minC, np.uint32) if minC is not None else None))
        if not asynchronous:
            self.synchronize()
        return flow, minC

    # ------------------------------------------------------------------ pyramid driver (pyramidal_sgm.m)
    def impyramid_reduce_dev(self, img, out):
        n, H, W = img.shape
        self._ck(self._l.fsgm_impyramid_reduce_dev(self._h, n, _dp(img), W, H, _dp(out)))

    def pyramidal_sgm_dev(self, I0, I1, mv, minC, opts=None, mvPyd=None):
        n, H, W = I0.shape
        self._ck(self._l.fsgm_pyramidal_sgm_dev(self._h, n, _dp(I0), _dp(I1), W, H, C.byref(opts) if opts is not None else None,
                                                _dp(mv), _dp(minC), _dp(mvPyd)))

    def pyramidal_sgm(self, I0, I1, opts=None, levels=False):
        """[mv, minC(, mvPyd)] = pyramidal_sgm(I0, I1, numPyd) on host arrays; mvPyd = list of per-level flows, finest first"""
        H, W = I0.shape
        o = opts if opts is not None else pyd_opts()
        ws, hs = pyramid_dims(W, H, o.numPyd)
        mv, minC = np.empty((2, H, W), np.float64), np.empty((H, W), np.uint32)
        flat = np.empty(sum(2 * w * h for w, h in zip(ws, hs)), np.float64) if levels else None
        self._ck(self._l.fsgm_pyramidal_sgm(self._h, _hp(I0, np.uint8), _hp(I1, np.uint8, (H, W)), W, H, C.byref(o),
                                            _hp(mv, np.float64), _hp(minC, np.uint32),
                                            _hp(flat, np.float64) if levels else None))
        if not levels:
            return mv, minC
        out, off = [], 0
        for w, h in zip(ws, hs):
            out.append(flat[off:off + 2 * w * h].reshape(2, h, w))
            off += 2 * w * h
        return mv, minC, out

    # ------------------------------------------------------------------ gateway 3: calc_cost_sgm_ng
    def calc_cost_sgm_ng(self, I1, I2, preMv=None, halfSearchWinSize=1, aggSize=2, subPixelRefine=0, P1=6, P2=32,
                         seed=1, rand_stream=None):
        """[minC, flow] = calc_cost_sgm_ng(...)  (calc_cost_sgm_ng.cpp:25, :484-527); the libc rand() state is an input"""
        H, W = I1.shape
        minC, flow = np.empty((H, W), np.uint32), np.empty((2, H, W), np.float64)
        o = NgOpts(int(seed), None)
        if rand_stream is not None:
            o.rand_stream = _hp(rand_stream, np.int32, (H * W * 8,)).value
        self._ck(self._l.fsgm_calc_cost_sgm_ng(
            self._h, _hp(I1, np.uint8), _hp(I2, np.uint8, (H, W)), W, H, _hp(preMv, np.float64) if preMv is not None else None,
            C.c_double(halfSearchWinSize), C.c_double(aggSize), int(subPixelRefine), int(P1), int(P2), C.byref(o),
            _hp(minC, np.uint32), _hp(flow, np.float64)))
        return minC, flow

    def calc_cost_sgm_ng_dev(self, I1, I2, P1, P2, minC, flow, seeds=None, Sp=None, Centries=None):
        n, H, W = I1.shape
        sd = (C.c_uint * n)(*[int(s) for s in seeds]) if seeds is not None else None
        self._ck(self._l.fsgm_calc_cost_sgm_ng_dev(self._h, n, _dp(I1), _dp(I2), W, H, int(P1), int(P2), sd, None,
                                                   _dp(minC), _dp(flow), _dp(Sp), _dp(Centries)))

    # ------------------------------------------------------------------ gateway 4: calc_pyd_cost_sgm_ng
    def calc_pyd_cost_sgm_ng(self, I1, I2, preMv, halfSearchWinSize, aggSize, subPixelRefine, P1, P2):
        """[minC, flow] = calc_pyd_cost_sgm_ng(...)  (calc_pyd_cost_sgm_ng.cpp:25, :448-523)"""
        H, W = I1.shape
        _, mvH, mvW = preMv.shape
        minC, flow = np.empty((H, W), np.uint32), np.empty((2, H, W), np.float64)
        self._ck(self._l.fsgm_calc_pyd_cost_sgm_ng(
            self._h, _hp(I1, np.uint8), _hp(I2, np.uint8, (H, W)), W, H, _hp(preMv, np.float64), mvW, mvH,
            int(halfSearchWinSize), int(aggSize), int(subPixelRefine), int(P1), int(P2), _hp(minC, np.uint32), _hp(flow, np.float64)))
        return minC, flow

    def calc_pyd_cost_sgm_ng_dev(self, I1, I2, preMv, r, aggSize, sub, P1, P2, minC, flow, Sp=None, cost=None, XY=None):
        n, H, W = I1.shape
        _, _, mvH, mvW = preMv.shape
        self._ck(self._l.fsgm_calc_pyd_cost_sgm_ng_dev(self._h, n, _dp(I1), _dp(I2), W, H, _dp(preMv), mvW, mvH, int(r), int(aggSize),
                                                       int(sub), int(P1), int(P2), _dp(minC), _dp(flow), _dp(Sp), _dp(cost), _dp(XY)))
