"""ctypes binding of the CPU oracle oracle/libba_ref.so (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

from orb_slam3_study_kr_b200.problem import (BAProblem, BAResult, CCamera, PoseBatch, PoseResult, Schedule,
                                              make_cameras)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libba_ref.so")
    src = os.path.join(_HERE, "ba_ref.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ba_ref_solve.restype = C.c_int
        _LIB.ba_ref_pose_opt_batch.restype = C.c_int
    return _LIB


def solve(problem: BAProblem, schedule: Schedule, with_counters: bool = False):
    L = lib()
    cp, keep1 = problem.to_c()
    cs, keep2 = schedule.to_c()
    res, cr, trace = BAResult.alloc(problem, schedule.max_trace)
    counters = (C.c_int64 * 4)()
    t0 = time.perf_counter()
    L.ba_ref_solve(C.byref(cp), C.byref(cs), C.byref(cr), counters)
    dt = time.perf_counter() - t0
    res.finish(cr, trace)
    if with_counters:
        return res, dict(seconds=dt, edge_linearisations=counters[0], edge_evaluations=counters[1],
                         lm_iterations=counters[2], lm_trials=counters[3])
    return res


def pose_opt_batch(batch: PoseBatch, with_time: bool = False):
    L = lib()
    cb, keep = batch.to_c()
    res, cr = PoseResult.alloc(batch)
    t0 = time.perf_counter()
    L.ba_ref_pose_opt_batch(C.byref(cb), C.byref(cr))
    dt = time.perf_counter() - t0
    return (res, dt) if with_time else res


def edge(kind: int, cam: dict, trl_qt, pose_qt, X, obs, unary: bool):
    """err(3), A(3x3), B(3x6), depth_positive of one edge at one state."""
    L = lib()
    cams = make_cameras([cam])
    err = np.zeros(3); A = np.zeros((3, 3)); B = np.zeros((3, 6)); dp = C.c_int(0)
    pose_qt = np.ascontiguousarray(pose_qt, np.float64); X = np.ascontiguousarray(X, np.float64)
    obs = np.ascontiguousarray(obs, np.float64)
    trl = None if trl_qt is None else np.ascontiguousarray(trl_qt, np.float64)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    L.ba_ref_edge(C.c_int(kind), C.byref(cams[0]), vp(trl), vp(pose_qt), vp(X), vp(obs), C.c_int(int(unary)),
                  vp(err), vp(A), vp(B), C.byref(dp))
    return err, A, B, bool(dp.value)


def oplus(pose_qt, update6):
    L = lib()
    pose_qt = np.ascontiguousarray(pose_qt, np.float64); u = np.ascontiguousarray(update6, np.float64)
    out = np.zeros(7)
    L.ba_ref_oplus(pose_qt.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def huber(e: float, delta: float):
    L = lib()
    r0, r1 = C.c_double(0), C.c_double(0)
    L.ba_ref_huber(C.c_double(e), C.c_double(delta), C.byref(r0), C.byref(r1))
    return r0.value, r1.value


def atan2f(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    L = lib()
    y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(y)
    L.ba_ref_atan2f(y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                    C.c_int64(y.size))
    return out


def dense_ldlt(H: np.ndarray, b: np.ndarray):
    L = lib()
    H = np.ascontiguousarray(H, np.float64); b = np.ascontiguousarray(b, np.float64)
    x = np.zeros_like(b)
    L.ba_ref_dense_ldlt.restype = C.c_int
    ok = L.ba_ref_dense_ldlt(C.c_int(len(b)), H.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                             x.ctypes.data_as(C.c_void_p))
    return bool(ok), x


def skyline_solve(H: np.ndarray, b: np.ndarray):
    L = lib()
    H = np.ascontiguousarray(H, np.float64); b = np.ascontiguousarray(b, np.float64)
    x = np.zeros_like(b)
    L.ba_ref_skyline_solve.restype = C.c_int
    ok = L.ba_ref_skyline_solve(C.c_int(len(b)), H.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                x.ctypes.data_as(C.c_void_p))
    return bool(ok), x
