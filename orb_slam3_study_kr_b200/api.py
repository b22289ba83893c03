"""ctypes binding of libbagpu.so (include/bagpu.h) and the host-side mirror of the reference's BA entry points.

The functions below keep the reference's names and call semantics (src/Optimizer.cc):

    local_bundle_adjustment      Optimizer::LocalBundleAdjustment(KeyFrame*, bool*, Map*, ...)      :1116-1498
    merge_bundle_adjustment      Optimizer::LocalBundleAdjustment(KeyFrame*, vector, vector, bool*)  :3506-3953
    global_bundle_adjustment     Optimizer::GlobalBundleAdjustemnt / BundleAdjustment                :53-390
    pose_optimization            Optimizer::PoseOptimization(Frame*) (batched over frames)           :815-1114

They take the flat problem (what the gather loops of Optimizer.cc produce), make ONE C-ABI call and return the
estimates plus the per-edge chi2 / depth flags the reference's classification loops read. There is no CPU path:
importing works anywhere, but creating a Context without the CUDA library or without a GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .problem import (BAProblem, BAResult, CTiming, PoseBatch, PoseResult, Schedule, schedule_global_ba,
                      schedule_local_ba, schedule_merge_ba)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbagpu.so")
_LIB = None

EXPORTS = ["bagpu_init", "bagpu_destroy", "bagpu_strerror", "bagpu_last_error", "bagpu_comm_unique_id",
           "bagpu_comm_init", "bagpu_pin_host", "bagpu_unpin_host", "bagpu_solve_ba", "bagpu_upload",
           "bagpu_solve_resident", "bagpu_download", "bagpu_reset_resident", "bagpu_update_estimates", "bagpu_pose_opt_batch", "bagpu_pose_upload",
           "bagpu_pose_solve_resident", "bagpu_get_timing", "bagpu_test_atan2f", "bagpu_test_solve", "bagpu_test_solve_parts", "bagpu_test_fp64_peak"]


class BagpuError(RuntimeError):
    pass


def load_library():
    """dlopen the in-tree libbagpu.so. Fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    lib_path = os.environ.get("BAGPU_LIB", LIB_PATH)       # development override: an alternative build of the same ABI
    if not os.path.exists(lib_path):
        raise BagpuError(f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         f"or `make -C {os.path.join(_HERE, 'csrc')}`. libbagpu has no CPU fallback.")
    L = C.CDLL(lib_path)
    L.bagpu_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.bagpu_init.restype = C.c_int
    L.bagpu_destroy.argtypes = [C.c_void_p]
    L.bagpu_destroy.restype = None
    L.bagpu_strerror.argtypes = [C.c_int]
    L.bagpu_strerror.restype = C.c_char_p
    L.bagpu_last_error.argtypes = [C.c_void_p]
    L.bagpu_last_error.restype = C.c_char_p
    L.bagpu_comm_unique_id.argtypes = [C.c_void_p]
    L.bagpu_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.bagpu_pin_host.argtypes = [C.c_void_p, C.c_size_t]
    L.bagpu_unpin_host.argtypes = [C.c_void_p]
    for name in ("bagpu_solve_ba",):
        getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bagpu_upload.argtypes = [C.c_void_p, C.c_void_p]
    L.bagpu_solve_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.bagpu_download.argtypes = [C.c_void_p, C.c_void_p]
    L.bagpu_reset_resident.argtypes = [C.c_void_p]
    L.bagpu_update_estimates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.bagpu_pose_opt_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.bagpu_pose_upload.argtypes = [C.c_void_p, C.c_void_p]
    L.bagpu_pose_solve_resident.argtypes = [C.c_void_p, C.c_void_p]
    L.bagpu_get_timing.argtypes = [C.c_void_p, C.c_void_p]
    L.bagpu_test_atan2f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    for n in EXPORTS:
        if n not in ("bagpu_destroy", "bagpu_strerror", "bagpu_last_error"):
            getattr(L, n).restype = C.c_int
    _LIB = L
    return L


class Context:
    """One bagpu_ctx: a stream plus a grow-only device arena. One per calling thread (SURVEY 8b Threading)."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.bagpu_init(device, C.byref(h))
        if rc != 0:
            raise BagpuError(f"bagpu_init failed: {self.lib.bagpu_strerror(rc).decode()} (libbagpu needs a CUDA device)")
        self.h = h
        self._pinned = []

    def close(self):
        if getattr(self, "h", None):
            for a in self._pinned:
                self.lib.bagpu_unpin_host(a.ctypes.data_as(C.c_void_p))
            self._pinned = []
            self.lib.bagpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc < 0:
            raise BagpuError(f"{self.lib.bagpu_strerror(rc).decode()}: {self.lib.bagpu_last_error(self.h).decode()}")
        return rc

    # -- multi-GPU plumbing (torch.distributed moves the 128-byte id; see bench.py)
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self._check(self.lib.bagpu_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, world: int, rank: int, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._check(self.lib.bagpu_comm_init(self.h, world, rank, buf))

    def pin(self, *arrays: np.ndarray):
        """Page-lock caller-owned gather buffers (a real adapter keeps persistent, registered gather arrays)."""
        for a in arrays:
            if a is not None and a.nbytes > 0:
                if self.lib.bagpu_pin_host(a.ctypes.data_as(C.c_void_p), a.nbytes) == 0:
                    self._pinned.append(a)

    def pin_problem(self, p: BAProblem):
        self.pin(p.pose_qt, p.points, p.obs_pose, p.obs_point, p.obs_cam, p.obs_rig, p.obs_kind, p.obs_flags,
                 p.obs_u, p.obs_v, p.obs_ur, p.obs_inv_sigma2)

    # -- BA
    def solve_ba(self, problem: BAProblem, schedule: Schedule, into=None) -> BAResult:
        """One bagpu_solve_ba call with host buffers. `into` = a (result, CResult, trace) triple from alloc_result() to
        reuse (and pin) the output buffers across calls, as a SLAM thread would."""
        cp, keep1 = problem.to_c()
        cs, keep2 = schedule.to_c()
        res, cr, trace = into if into is not None else BAResult.alloc(problem, schedule.max_trace)
        self._check(self.lib.bagpu_solve_ba(self.h, C.byref(cp), C.byref(cs), C.byref(cr)))
        return res.finish(cr, trace)

    def alloc_result(self, problem: BAProblem, schedule: Schedule, pinned: bool = True):
        """Reusable output buffers for solve_ba / solve_resident (page-locked when `pinned`)."""
        triple = BAResult.alloc(problem, schedule.max_trace)
        if pinned:
            r = triple[0]
            self.pin(r.pose_qt, r.points, r.edge_chi2, r.edge_depth_pos, r.edge_level)
        return triple

    def upload(self, problem: BAProblem):
        cp, keep = problem.to_c()
        self._check(self.lib.bagpu_upload(self.h, C.byref(cp)))
        self._problem = problem

    def solve_resident(self, schedule: Schedule, download: bool = True, into=None) -> Optional[BAResult]:
        cs, keep = schedule.to_c()
        if not download:
            self._check(self.lib.bagpu_solve_resident(self.h, C.byref(cs), None))
            return None
        res, cr, trace = into if into is not None else BAResult.alloc(self._problem, schedule.max_trace)
        self._check(self.lib.bagpu_solve_resident(self.h, C.byref(cs), C.byref(cr)))
        return res.finish(cr, trace)

    def reset_resident(self):
        self._check(self.lib.bagpu_reset_resident(self.h))

    def update_estimates(self, pose_qt=None, points=None):
        """New estimates for the resident map (same structure): the upload's plan is kept (bagpu_update_estimates)."""
        pq = None if pose_qt is None else np.ascontiguousarray(pose_qt, np.float64)
        pt = None if points is None else np.ascontiguousarray(points, np.float64)
        self._check(self.lib.bagpu_update_estimates(self.h, None if pq is None else pq.ctypes.data, None if pt is None else pt.ctypes.data))

    # -- PoseOptimization
    def pose_opt_batch(self, batch: PoseBatch) -> PoseResult:
        cb, keep = batch.to_c()
        res, cr = PoseResult.alloc(batch)
        self._check(self.lib.bagpu_pose_opt_batch(self.h, C.byref(cb), C.byref(cr)))
        return res

    def pose_upload(self, batch: PoseBatch):
        cb, keep = batch.to_c()
        self._check(self.lib.bagpu_pose_upload(self.h, C.byref(cb)))
        self._batch = batch

    def pose_solve_resident(self, download: bool = True) -> Optional[PoseResult]:
        if not download:
            self._check(self.lib.bagpu_pose_solve_resident(self.h, None))
            return None
        res, cr = PoseResult.alloc(self._batch)
        self._check(self.lib.bagpu_pose_solve_resident(self.h, C.byref(cr)))
        return res

    def timing(self) -> dict:
        t = CTiming()
        self._check(self.lib.bagpu_get_timing(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in CTiming._fields_}

    def test_solve(self, A: np.ndarray, b: np.ndarray, col_end: np.ndarray, lam: float = 0.0, parts: int = 1):
        """(A + lam I) x = b with the production reduced-system solver (unit-test hook); parts >= 3: the partitioned solver."""
        A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64)
        ce = np.ascontiguousarray(col_end, np.int32)
        x = np.zeros_like(b); f = C.c_int(0)
        self.lib.bagpu_test_solve_parts.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        self._check(self.lib.bagpu_test_solve_parts(self.h, len(b), ce.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p),
                                                    b.ctypes.data_as(C.c_void_p), lam, parts, x.ctypes.data_as(C.c_void_p), C.byref(f)))
        return x, bool(f.value)

    def fp64_peak(self) -> dict:
        """Measured FP64 throughput of the device in TFLOP/s (DFMA and FP64 MMA probes)."""
        a, b = C.c_double(0), C.c_double(0)
        self.lib.bagpu_test_fp64_peak.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self.lib.bagpu_test_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return {"dfma_tflops": a.value, "dmma_tflops": b.value}

    def device_atan2f(self, y: np.ndarray, x: np.ndarray) -> np.ndarray:
        y = np.ascontiguousarray(y, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros_like(y)
        self._check(self.lib.bagpu_test_atan2f(self.h, y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
                                               out.ctypes.data_as(C.c_void_p), y.size))
        return out


# ------------------------------------------------------------------------------------------------
# Host-side mirror of the reference's entry points (same names, argument meaning and error behaviour).
_DEFAULT: Optional[Context] = None


def default_context() -> Context:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Context()
    return _DEFAULT


def local_bundle_adjustment(problem: BAProblem, stop_flag: Optional[np.ndarray] = None, inertial_map: bool = False,
                            ctx: Optional[Context] = None) -> BAResult:
    """LocalMapping's LBA: optimize(10) with Huber sqrt(5.991)/sqrt(7.815); the caller erases
    `result.outliers(problem)` observations and writes poses/points back (src/Optimizer.cc:1410-1497)."""
    s = schedule_local_ba(inertial_map)
    s.stop_flag = stop_flag
    # stop flag already set: the reference returns before optimising (Optimizer.cc:1406-1408). The library does the same:
    # status BAGPU_STOPPED, empty trace, the (normalised) input estimates back -- no exception, as in the reference.
    return (ctx or default_context()).solve_ba(problem, s)


def merge_bundle_adjustment(problem: BAProblem, stop_flag: Optional[np.ndarray] = None,
                            ctx: Optional[Context] = None) -> BAResult:
    """Merge/welding LBA: optimize(5), gate to level 1, drop kernels, optimize(10) (src/Optimizer.cc:3731-3780)."""
    s = schedule_merge_ba()
    s.stop_flag = stop_flag
    return (ctx or default_context()).solve_ba(problem, s)


def global_bundle_adjustment(problem: BAProblem, n_iterations: int = 5, stop_flag: Optional[np.ndarray] = None,
                             ctx: Optional[Context] = None) -> BAResult:
    """GlobalBundleAdjustemnt(pMap, nIterations, pbStopFlag, nLoopKF, bRobust): bRobust lives in problem.obs_flags."""
    s = schedule_global_ba(n_iterations)
    s.stop_flag = stop_flag
    return (ctx or default_context()).solve_ba(problem, s)


def pose_optimization(batch: PoseBatch, ctx: Optional[Context] = None) -> PoseResult:
    """PoseOptimization over a batch of independent frames; result.n_inliers[f] is the function's return value."""
    return (ctx or default_context()).pose_opt_batch(batch)
