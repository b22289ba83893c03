"""Flat BA problem containers and their ctypes view of include/bagpu.h.

This is the data that the reference's Optimizer.cc puts into g2o::SparseOptimizer
(src/Optimizer.cc:115-270 for global BA, :1213-1403 for local BA, :827-993 for
PoseOptimization), flattened: poses in vertex-id order, points in vertex-id order,
observations in edge insertion order.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

EDGE_MONO, EDGE_STEREO, EDGE_BODY = 0, 1, 2
CAM_PINHOLE, CAM_KB8 = 0, 1
FLAG_ROBUST = 1
GATE_NONE, GATE_LBA, GATE_POSE = 0, 1, 2
SOLVER_AUTO, SOLVER_CHOLESKY, SOLVER_PCG = 0, 1, 2

STATUS_NAMES = {0: "OK", 1: "TERMINATE_TRIALS", 2: "TERMINATE_NBAD", 3: "STOPPED"}

# Golden constants the reference source fixes:
#   const float thHuberMono = sqrt(5.991)   src/Optimizer.cc:1275   (LBA, PoseOptimization :853)
#   const float thHuber2D   = sqrt(5.99)    src/Optimizer.cc:132    (GBA), :3628 (merge LBA)
#   const float thHuber3D   = sqrt(7.815)   src/Optimizer.cc:133,1276,3629
DELTA_MONO_LBA = float(np.float32(np.sqrt(5.991)))
DELTA_MONO_GBA = float(np.float32(np.sqrt(5.99)))
DELTA_STEREO = float(np.float32(np.sqrt(7.815)))
GATE_MONO = 5.991
GATE_STEREO = 7.815


def inv_level_sigma2(n_levels: int = 8, scale_factor: float = 1.2) -> np.ndarray:
    """mvInvLevelSigma2, float recurrence of src/ORBextractor.cc:414-429."""
    sf = np.float32(scale_factor)
    scale = np.zeros(n_levels, np.float32)
    sigma2 = np.zeros(n_levels, np.float32)
    scale[0] = np.float32(1.0)
    sigma2[0] = np.float32(1.0)
    for i in range(1, n_levels):
        scale[i] = np.float32(scale[i - 1] * sf)
        sigma2[i] = np.float32(scale[i] * scale[i])
    return (np.float32(1.0) / sigma2).astype(np.float32)


# ----------------------------------------------------------------------------- ctypes structs
class CCamera(C.Structure):
    _fields_ = [("type", C.c_int32), ("p", C.c_float * 8), ("bf", C.c_float)]


class CRig(C.Structure):
    _fields_ = [("qt", C.c_double * 7)]


class CProblem(C.Structure):
    _fields_ = [
        ("n_poses", C.c_int32), ("pose_qt", C.c_void_p), ("pose_fixed", C.c_void_p),
        ("n_points", C.c_int32), ("points", C.c_void_p),
        ("n_cameras", C.c_int32), ("cameras", C.c_void_p),
        ("n_rigs", C.c_int32), ("rigs", C.c_void_p),
        ("n_obs", C.c_int64),
        ("obs_pose", C.c_void_p), ("obs_point", C.c_void_p), ("obs_cam", C.c_void_p),
        ("obs_rig", C.c_void_p), ("obs_kind", C.c_void_p), ("obs_flags", C.c_void_p),
        ("obs_u", C.c_void_p), ("obs_v", C.c_void_p), ("obs_ur", C.c_void_p),
        ("obs_inv_sigma2", C.c_void_p),
    ]


class CRound(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("gate_after", C.c_int32), ("gate_mono", C.c_double),
                ("gate_stereo", C.c_double), ("drop_kernel_after", C.c_int32), ("reset_pose", C.c_int32)]


class CSchedule(C.Structure):
    _fields_ = [("n_rounds", C.c_int32), ("rounds", C.c_void_p), ("delta_mono", C.c_double),
                ("delta_stereo", C.c_double), ("lambda_init", C.c_double), ("stop_flag", C.c_void_p),
                ("linear_solver", C.c_int32), ("max_trace", C.c_int32),
                ("pcg_tolerance", C.c_double), ("pcg_max_iterations", C.c_int32)]


class CTrace(C.Structure):
    _fields_ = [("round", C.c_int32), ("iteration", C.c_int32), ("chi2_before", C.c_double),
                ("chi2_after", C.c_double), ("lambda_", C.c_double), ("trials", C.c_int32), ("status", C.c_int32),
                ("active_edges", C.c_int64), ("linearise_schur_us", C.c_double), ("linear_solve_us", C.c_double),
                ("update_us", C.c_double), ("iteration_us", C.c_double)]


class CResult(C.Structure):
    _fields_ = [("pose_qt", C.c_void_p), ("points", C.c_void_p), ("edge_chi2", C.c_void_p),
                ("edge_depth_pos", C.c_void_p), ("edge_level", C.c_void_p), ("trace", C.c_void_p),
                ("n_trace", C.c_int32), ("status", C.c_int32)]


class CPoseBatch(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int32), ("pose_qt", C.c_void_p), ("frame_ptr", C.c_void_p),
        ("n_cameras", C.c_int32), ("cameras", C.c_void_p), ("n_rigs", C.c_int32), ("rigs", C.c_void_p),
        ("n_obs", C.c_int64), ("xw", C.c_void_p), ("obs_cam", C.c_void_p), ("obs_rig", C.c_void_p),
        ("obs_kind", C.c_void_p), ("obs_u", C.c_void_p), ("obs_v", C.c_void_p), ("obs_ur", C.c_void_p),
        ("obs_inv_sigma2", C.c_void_p), ("delta_mono", C.c_double), ("delta_stereo", C.c_double),
        ("gate_mono", C.c_float), ("gate_stereo", C.c_float),
    ]


class CPoseResult(C.Structure):
    _fields_ = [("pose_qt", C.c_void_p), ("outlier", C.c_void_p), ("n_inliers", C.c_void_p), ("final_chi2", C.c_void_p)]


class CTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("solve_ms", C.c_double), ("d2h_ms", C.c_double),
                ("build_ms", C.c_double), ("linsolve_ms", C.c_double), ("update_ms", C.c_double),
                ("build_launches", C.c_int64), ("update_launches", C.c_int64), ("linsolve_launches", C.c_int64),
                ("total_launches", C.c_int64), ("lm_iterations", C.c_int64), ("lm_trials", C.c_int64),
                ("edge_linearisations", C.c_int64), ("edge_evaluations", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("pcg_iterations", C.c_int32), ("schur_blocks", C.c_int32),
                ("solve_retries", C.c_int32), ("solver_parts", C.c_int32)]


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_cameras(cams: List[dict]):
    arr = (CCamera * max(1, len(cams)))()
    for i, c in enumerate(cams):
        arr[i].type = int(c["type"])
        p = list(c["p"]) + [0.0] * (8 - len(c["p"]))
        for k in range(8):
            arr[i].p[k] = float(np.float32(p[k]))
        arr[i].bf = float(np.float32(c.get("bf", 0.0)))
    return arr


def make_rigs(rigs: np.ndarray):
    n = 0 if rigs is None else len(rigs)
    arr = (CRig * max(1, n))()
    for i in range(n):
        for k in range(7):
            arr[i].qt[k] = float(rigs[i][k])
    return arr


# ----------------------------------------------------------------------------- containers
@dataclass
class Round:
    iterations: int
    gate_after: int = GATE_NONE
    gate_mono: float = GATE_MONO
    gate_stereo: float = GATE_STEREO
    drop_kernel_after: bool = False
    reset_pose: bool = False


@dataclass
class Schedule:
    rounds: List[Round]
    delta_mono: float = DELTA_MONO_LBA
    delta_stereo: float = DELTA_STEREO
    lambda_init: float = 0.0
    linear_solver: int = SOLVER_AUTO
    max_trace: int = 256
    stop_flag: Optional[np.ndarray] = None      # uint8[1], written by another thread
    pcg_tolerance: float = 0.0                  # SOLVER_PCG only (0: library default 1e-10)
    pcg_max_iterations: int = 0

    def total_iterations(self) -> int:
        return sum(r.iterations for r in self.rounds)

    def to_c(self):
        rounds = (CRound * len(self.rounds))()
        for i, r in enumerate(self.rounds):
            rounds[i] = CRound(r.iterations, r.gate_after, r.gate_mono, r.gate_stereo,
                               int(r.drop_kernel_after), int(r.reset_pose))
        s = CSchedule(len(self.rounds), C.cast(rounds, C.c_void_p), self.delta_mono, self.delta_stereo,
                      self.lambda_init, _ptr(self.stop_flag), self.linear_solver, self.max_trace,
                      self.pcg_tolerance, self.pcg_max_iterations)
        return s, rounds


# Schedules of the four entry points -------------------------------------------------
def schedule_local_ba(inertial: bool = False) -> Schedule:
    """Optimizer::LocalBundleAdjustment(KeyFrame*,bool*,Map*,...): one optimize(10), src/Optimizer.cc:1410-1411."""
    return Schedule([Round(10)], DELTA_MONO_LBA, DELTA_STEREO, 100.0 if inertial else 0.0)


def schedule_merge_ba() -> Schedule:
    """Merge LBA: optimize(5) -> gate -> kernels off -> optimize(10), src/Optimizer.cc:3731-3780."""
    return Schedule([Round(5, GATE_LBA, GATE_MONO, GATE_STEREO, True), Round(10)], DELTA_MONO_GBA, DELTA_STEREO)


def schedule_global_ba(iterations: int = 10) -> Schedule:
    """Optimizer::BundleAdjustment: one optimize(nIterations), src/Optimizer.cc:278-280. Robustness is per-edge flags."""
    return Schedule([Round(iterations)], DELTA_MONO_GBA, DELTA_STEREO)


@dataclass
class BAProblem:
    pose_qt: np.ndarray            # [Nt,7] f64: tx ty tz qx qy qz qw
    pose_fixed: np.ndarray         # [Nt] u8
    points: np.ndarray             # [Np,3] f64
    cameras: List[dict]
    rigs: Optional[np.ndarray]     # [Nr,7] f64 or None
    obs_pose: np.ndarray           # i32
    obs_point: np.ndarray          # i32
    obs_cam: np.ndarray            # i16
    obs_rig: np.ndarray            # i16
    obs_kind: np.ndarray           # u8
    obs_flags: np.ndarray          # u8
    obs_u: np.ndarray              # f64
    obs_v: np.ndarray
    obs_ur: np.ndarray
    obs_inv_sigma2: np.ndarray
    name: str = ""
    truth: dict = field(default_factory=dict)

    def __post_init__(self):
        self.pose_qt = np.ascontiguousarray(self.pose_qt, np.float64)
        self.pose_fixed = np.ascontiguousarray(self.pose_fixed, np.uint8)
        self.points = np.ascontiguousarray(self.points, np.float64)
        self.obs_pose = np.ascontiguousarray(self.obs_pose, np.int32)
        self.obs_point = np.ascontiguousarray(self.obs_point, np.int32)
        self.obs_cam = np.ascontiguousarray(self.obs_cam, np.int16)
        self.obs_rig = np.ascontiguousarray(self.obs_rig, np.int16)
        self.obs_kind = np.ascontiguousarray(self.obs_kind, np.uint8)
        self.obs_flags = np.ascontiguousarray(self.obs_flags, np.uint8)
        for k in ("obs_u", "obs_v", "obs_ur", "obs_inv_sigma2"):
            setattr(self, k, np.ascontiguousarray(getattr(self, k), np.float64))

    @property
    def n_poses(self): return int(self.pose_qt.shape[0])
    @property
    def n_points(self): return int(self.points.shape[0])
    @property
    def n_obs(self): return int(self.obs_pose.shape[0])
    @property
    def n_free(self): return int((self.pose_fixed == 0).sum())

    def input_bytes(self) -> int:
        return int(self.pose_qt.nbytes + self.pose_fixed.nbytes + self.points.nbytes + self.obs_pose.nbytes +
                   self.obs_point.nbytes + self.obs_cam.nbytes + self.obs_rig.nbytes + self.obs_kind.nbytes +
                   self.obs_flags.nbytes + self.obs_u.nbytes + self.obs_v.nbytes + self.obs_ur.nbytes +
                   self.obs_inv_sigma2.nbytes)

    def to_c(self):
        cams = make_cameras(self.cameras)
        rigs = make_rigs(self.rigs)
        n_rigs = 0 if self.rigs is None else len(self.rigs)
        p = CProblem(self.n_poses, _ptr(self.pose_qt), _ptr(self.pose_fixed), self.n_points, _ptr(self.points),
                     len(self.cameras), C.cast(cams, C.c_void_p), n_rigs, C.cast(rigs, C.c_void_p), self.n_obs,
                     _ptr(self.obs_pose), _ptr(self.obs_point), _ptr(self.obs_cam), _ptr(self.obs_rig),
                     _ptr(self.obs_kind), _ptr(self.obs_flags), _ptr(self.obs_u), _ptr(self.obs_v),
                     _ptr(self.obs_ur), _ptr(self.obs_inv_sigma2))
        return p, (cams, rigs)

    def shard_by_landmark(self, rank: int, world: int) -> "BAProblem":
        """Landmark shard for multi-GPU global BA: contiguous point range, all of its
        observations, all poses/cameras replicated. Observation order within the shard is preserved."""
        Np = self.n_points
        lo = (Np * rank) // world
        hi = (Np * (rank + 1)) // world
        sel = (self.obs_point >= lo) & (self.obs_point < hi)
        return BAProblem(self.pose_qt, self.pose_fixed, self.points[lo:hi], self.cameras, self.rigs,
                         self.obs_pose[sel], self.obs_point[sel] - lo, self.obs_cam[sel], self.obs_rig[sel],
                         self.obs_kind[sel], self.obs_flags[sel], self.obs_u[sel], self.obs_v[sel],
                         self.obs_ur[sel], self.obs_inv_sigma2[sel], name=f"{self.name}[shard {rank}/{world}]",
                         truth={"point_range": (lo, hi), "obs_mask": sel})


@dataclass
class BAResult:
    pose_qt: np.ndarray
    points: np.ndarray
    edge_chi2: np.ndarray
    edge_depth_pos: np.ndarray
    edge_level: np.ndarray
    trace: List[dict]
    status: int

    @staticmethod
    def alloc(problem: BAProblem, max_trace: int):
        r = BAResult(np.zeros((problem.n_poses, 7)), np.zeros((problem.n_points, 3)), np.zeros(problem.n_obs),
                     np.zeros(problem.n_obs, np.uint8), np.zeros(problem.n_obs, np.uint8), [], 0)
        trace = (CTrace * max(1, max_trace))()
        c = CResult(_ptr(r.pose_qt), _ptr(r.points), _ptr(r.edge_chi2), _ptr(r.edge_depth_pos),
                    _ptr(r.edge_level), C.cast(trace, C.c_void_p), 0, 0)
        return r, c, trace

    def finish(self, c: "CResult", trace):
        self.status = int(c.status)
        self.trace = [dict(round=t.round, iteration=t.iteration, chi2_before=t.chi2_before, chi2_after=t.chi2_after,
                           lambda_=t.lambda_, trials=t.trials, status=t.status, active_edges=t.active_edges,
                           linearise_schur_us=t.linearise_schur_us, linear_solve_us=t.linear_solve_us, update_us=t.update_us,
                           iteration_us=t.iteration_us) for t in trace[: c.n_trace]]
        return self

    def outliers(self, problem: BAProblem, gate_mono=GATE_MONO, gate_stereo=GATE_STEREO) -> np.ndarray:
        """The reference's final classification: chi2 > gate || !isDepthPositive (src/Optimizer.cc:1416-1460)."""
        th = np.where(problem.obs_kind == EDGE_STEREO, gate_stereo, gate_mono)
        return (self.edge_chi2 > th) | (self.edge_depth_pos == 0)


@dataclass
class PoseBatch:
    pose_qt: np.ndarray            # [F,7]
    frame_ptr: np.ndarray          # [F+1] i64
    cameras: List[dict]
    rigs: Optional[np.ndarray]
    xw: np.ndarray                 # [Ne,3]
    obs_cam: np.ndarray
    obs_rig: np.ndarray
    obs_kind: np.ndarray
    obs_u: np.ndarray
    obs_v: np.ndarray
    obs_ur: np.ndarray
    obs_inv_sigma2: np.ndarray
    delta_mono: float = DELTA_MONO_LBA
    delta_stereo: float = DELTA_STEREO
    gate_mono: float = 5.991
    gate_stereo: float = 7.815
    truth: dict = field(default_factory=dict)

    def __post_init__(self):
        self.pose_qt = np.ascontiguousarray(self.pose_qt, np.float64)
        self.frame_ptr = np.ascontiguousarray(self.frame_ptr, np.int64)
        self.xw = np.ascontiguousarray(self.xw, np.float64)
        self.obs_cam = np.ascontiguousarray(self.obs_cam, np.int16)
        self.obs_rig = np.ascontiguousarray(self.obs_rig, np.int16)
        self.obs_kind = np.ascontiguousarray(self.obs_kind, np.uint8)
        for k in ("obs_u", "obs_v", "obs_ur", "obs_inv_sigma2"):
            setattr(self, k, np.ascontiguousarray(getattr(self, k), np.float64))

    @property
    def n_frames(self): return int(self.pose_qt.shape[0])
    @property
    def n_obs(self): return int(self.xw.shape[0])

    def input_bytes(self) -> int:
        return int(self.pose_qt.nbytes + self.frame_ptr.nbytes + self.xw.nbytes + self.obs_cam.nbytes +
                   self.obs_rig.nbytes + self.obs_kind.nbytes + self.obs_u.nbytes + self.obs_v.nbytes +
                   self.obs_ur.nbytes + self.obs_inv_sigma2.nbytes)

    def to_c(self):
        cams = make_cameras(self.cameras)
        rigs = make_rigs(self.rigs)
        n_rigs = 0 if self.rigs is None else len(self.rigs)
        b = CPoseBatch(self.n_frames, _ptr(self.pose_qt), _ptr(self.frame_ptr), len(self.cameras),
                       C.cast(cams, C.c_void_p), n_rigs, C.cast(rigs, C.c_void_p), self.n_obs, _ptr(self.xw),
                       _ptr(self.obs_cam), _ptr(self.obs_rig), _ptr(self.obs_kind), _ptr(self.obs_u), _ptr(self.obs_v),
                       _ptr(self.obs_ur), _ptr(self.obs_inv_sigma2), self.delta_mono, self.delta_stereo,
                       float(np.float32(self.gate_mono)), float(np.float32(self.gate_stereo)))
        return b, (cams, rigs)


@dataclass
class PoseResult:
    pose_qt: np.ndarray
    outlier: np.ndarray
    n_inliers: np.ndarray
    final_chi2: np.ndarray

    @staticmethod
    def alloc(batch: PoseBatch):
        r = PoseResult(np.zeros((batch.n_frames, 7)), np.zeros(batch.n_obs, np.uint8),
                       np.zeros(batch.n_frames, np.int32), np.zeros(batch.n_frames))
        c = CPoseResult(_ptr(r.pose_qt), _ptr(r.outlier), _ptr(r.n_inliers), _ptr(r.final_chi2))
        return r, c
