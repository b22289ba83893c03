#!/usr/bin/env python
"""bench.py -- global bundle adjustment on N B200s, one JSON line.

  python bench.py --gpus N --steps K --warmup W            libbagpu arm
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle restatement, 1 thread)

Workload (config.workload), the same map at every N (STRONG scaling):
  c5 (default)  BASELINE.json config 5: 5000 keyframes (1 fixed), 2 000 000 landmarks, ~20 M stereo/mono observations, EuRoC
                intrinsics, seed 5, non-robust (LoopClosing's GBA call, src/LoopClosing.cc:2289), optimize(20). It fits one GPU,
                so N=1 solves the whole map; at N>1 rank r owns the contiguous landmark range [r Np/N, (r+1) Np/N) (MapPoint ids
                grow with the keyframe that created them, so a range is covisibility-local) and all of its observations.
  c4            BASELINE.json config 4: 500 keyframes, 200 000 landmarks, ~2 M observations, seed 4, sharded the same way.
The other workload is measured after the headline and reported under "also" (device-resident and end to end), so one run
carries both strong-scaling curves. A step = one Optimizer::GlobalBundleAdjustemnt call = optimize(20) from the initial estimates.

  value   edge passes per second (edge linearisations + edge evaluations, summed over ranks) with the map resident in
          HBM, timed with CUDA events on the library's stream, max over ranks.
  e2e     the same metric through the reference-facing C-ABI call bagpu_solve_ba with HOST buffers: H2D of the whole
          problem, solve, D2H of poses / points / per-edge chi2 / flags inside the timed region.
  roofline  the linearise+Schur pass (stage_kernel + tile_diag_kernel + pair_tile_mma_kernel), the dominant kernel group of the library stream:
          algorithmic bytes per pass / mean pass duration (CUDA events around every pass in the timed region, on the
          library's stream) against the measured HBM copy bandwidth; roofline_fp64 reports the same pass against the FP64
          throughput measured on the device (the Schur products are FP64 work; they run on the FP64 tensor pipe,
          mma.sync.m8n8k4.f64). On config 4 (one GPU) the two-front band Cholesky runs BESIDE the pass on a second stream, so the
          `kernels` shares overlap and sum to more than 1; on config 5 the 13-front solver runs after it.
  parity_check  before the timed region every run solves a small map (config 4 at 1/10 size, merge schedule: two rounds and
          a gate) sharded over the N ranks and compares rank 0's result with the CPU oracle (1e-6 on chi2 per iteration
          and on the estimates, identical trial counts and edge levels): the multi-GPU path is checked where it is timed.
  local_ba  secondary lines (1 GPU): configs 1-3 through bagpu_solve_ba and the 1000-frame PoseOptimization batch, end to end.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "global BA edge passes/s (LM linearisations + evaluations per second; LM iters/s alongside)"
N_ITER = 20
WORKLOADS = {"c5": (5, "C5 global BA (BASELINE config 5): 5000 KFs (1 fixed), 2M landmarks, %d observations, seed 5"),
             "c4": (4, "C4 global BA (BASELINE config 4): 500 KFs (1 fixed), 200k landmarks, %d observations, seed 4")}


def workload_name(which: str, n_obs_total: int) -> str:
    return (WORKLOADS[which][1] % n_obs_total) + ", EuRoC stereo+mono, non-robust, optimize(%d), strong-scaled over contiguous landmark ranges" % N_ITER


# ----------------------------------------------------------------------------- distributed helpers (also used by tests)
def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def broadcast_bytes(data, n: int) -> bytes:
    """Rank 0's `n` bytes to every rank over whatever torch.distributed backend is up (nccl -> cuda, gloo -> cpu)."""
    import torch
    dist = _dist()
    if dist is None:
        return data
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(n, dtype=torch.uint8, device=dev)
    if dist.get_rank() == 0:
        t.copy_(torch.tensor(list(data), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def max_over_ranks(x: float) -> float:
    import torch
    dist = _dist()
    if dist is None:
        return float(x)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float) -> float:
    import torch
    dist = _dist()
    if dist is None:
        return float(x)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch
    dist = _dist()
    if dist is not None:
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- workload
_FULL = {}


def full_map(which: str, robust: bool = False):
    from orb_slam3_study_kr_b200 import synthetic
    key = (which, robust)
    if key not in _FULL:
        _FULL[key] = synthetic.config(WORKLOADS[which][0], robust=robust)
    return _FULL[key]


def make_workload(which: str, rank: int, world: int, robust: bool = False):
    """Rank `rank`'s landmark range of the full map (the whole map at world == 1). Every rank generates the same seeded map
    and keeps its slice, so the union over ranks is exactly the BASELINE config at every N."""
    full = full_map(which, robust)
    return (full if world == 1 else full.shard_by_landmark(rank, world)), full.n_obs


def pass_algorithmic(p, n_free: int, schur_blocks: int):
    """SURVEY.md 8(d). Bytes: per observation 40 B (mono) / 48 B (stereo) read; per landmark 24 B read; the reduced camera
    system written once: 288 B per upper 6x6 block INSIDE the stored band (bagpu_timing.schur_blocks) + 48 B per camera.
    FP64 work of the Schur products: sum over landmarks of 54 k + 108 k(k+1)/2 + 27 multiply-adds (k = track length)."""
    from orb_slam3_study_kr_b200.problem import EDGE_STEREO
    n_st = int((p.obs_kind == EDGE_STEREO).sum())
    nbytes = 48.0 * n_st + 40.0 * (p.n_obs - n_st) + 24.0 * p.n_points + 288.0 * schur_blocks + 48.0 * n_free
    k = np.bincount(p.obs_point, minlength=p.n_points).astype(np.float64)
    k = k[k > 0]
    mac = float((54.0 * k + 108.0 * k * (k + 1) / 2 + 27.0).sum())
    return nbytes, 2.0 * mac


def peaks():
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        d = json.load(open(f))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- CPU reference arm
def run_reference(args, rank: int, world: int):
    """The reference's BA is single-threaded by construction (Thirdparty/g2o/CMakeLists.txt:48 G2O_USE_OPENMP OFF) and
    cannot be built here (no Eigen3/OpenCV), so this arm times the oracle restatement on one host core. Each step is a
    bounded sample: `ref_iters` LM iterations of the same (whole) map; the rate does not depend on N."""
    if rank != 0:
        return
    from orb_slam3_study_kr_b200.problem import schedule_global_ba
    from oracle import ba_ref
    which = args.workload
    p, n_obs_total = make_workload(which, 0, 1)
    ref_iters = args.ref_iters if args.ref_iters > 0 else (1 if which == "c5" else 4)
    s = schedule_global_ba(ref_iters)
    if which != "c5":
        for _ in range(min(args.warmup, 1)):
            ba_ref.solve(p, schedule_global_ba(1))
    tot_s, passes, iters = 0.0, 0, 0
    for _ in range(args.steps):
        _, c = ba_ref.solve(p, s, True)
        tot_s += c["seconds"]
        passes += c["edge_linearisations"] + c["edge_evaluations"]
        iters += c["lm_iterations"]
    v = passes / tot_s
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "edge passes/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "lm_iters_per_s": iters / tot_s,
            "config": {"workload": workload_name(which, n_obs_total),
                       "sample": "%d of %d LM iterations per step on the whole map, one host thread (the rate does not depend on N)" % (ref_iters, N_ITER)},
            "cpu_baseline": {"value": v, "unit": "edge passes/s", "cores": 1, "kind": "port",
                             "sample": "%d steps x %d LM iterations of the %s map, oracle/ba_ref.cpp -O3 -march=x86-64-v3, 1 thread "
                                       "(reference g2o is built without OpenMP)" % (args.steps, ref_iters, which.upper())},
            "e2e": {"value": v, "unit": "edge passes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- secondary lines: local BA, pose batch (1 GPU)
def local_ba_lines(ctx, l2_flush, reps: int = 5):
    """BASELINE.json configs 1-3 through bagpu_solve_ba / bagpu_pose_opt_batch with host buffers (end to end, wall clock):
    Optimizer::LocalBundleAdjustment (LocalMapping schedule: optimize(10), robust) and batched Optimizer::PoseOptimization."""
    from orb_slam3_study_kr_b200 import synthetic
    from orb_slam3_study_kr_b200.problem import schedule_local_ba
    out = {}
    s = schedule_local_ba()
    for cfg in (1, 2, 3):
        p = synthetic.config(cfg)
        ctx.pin_problem(p)
        buf = ctx.alloc_result(p, s)
        ctx.solve_ba(p, s, into=buf)
        ts, passes, iters = [], 0, 0
        for _ in range(reps):
            l2_flush()
            t0 = time.perf_counter()
            ctx.solve_ba(p, s, into=buf)
            ts.append(time.perf_counter() - t0)
            t = ctx.timing()
            passes, iters = t["edge_linearisations"] + t["edge_evaluations"], t["lm_iterations"]
        ms = 1e3 * float(np.median(ts))
        out["C%d" % cfg] = {"observations": int(p.n_obs), "free_poses": int(p.n_free), "points": int(p.n_points), "e2e_ms_per_call": ms,
                            "lm_iterations": int(iters), "lm_iters_per_s": iters / (ms * 1e-3), "edge_passes_per_s": passes / (ms * 1e-3)}
    b = synthetic.make_pose_batch(n_frames=1000, n_matches=300)
    ctx.pose_opt_batch(b)
    ts = []
    for _ in range(reps):
        l2_flush()
        t0 = time.perf_counter()
        ctx.pose_opt_batch(b)
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.median(ts))
    out["pose_opt_batch"] = {"frames": 1000, "edges": int(b.n_obs), "e2e_ms_per_call": ms, "frames_per_s": 1000 / (ms * 1e-3)}
    return out


# ----------------------------------------------------------------------------- GPU arm
def parity_check(ctx, rank: int, world: int):
    """A small map (config 4 at 1/10 size: 50 keyframes, 20k landmarks, ~200k observations; merge schedule = two rounds with a
    gate, so every collective runs) sharded over the N ranks exactly like the timed workload, against the CPU oracle on rank 0."""
    from orb_slam3_study_kr_b200 import synthetic
    from orb_slam3_study_kr_b200.problem import schedule_merge_ba
    full = synthetic.config(4, scale=0.1, robust=True)
    mine = full if world == 1 else full.shard_by_landmark(rank, world)
    s = schedule_merge_ba()
    got = ctx.solve_ba(mine, s)
    out = None
    if rank == 0:
        from oracle import ba_ref
        ref = ba_ref.solve(full, s)
        lo, hi = (0, full.n_points) if world == 1 else mine.truth["point_range"]
        sel = slice(None) if world == 1 else mine.truth["obs_mask"]
        same_len = len(got.trace) == len(ref.trace)
        rel = max([abs(a["chi2_after"] - b["chi2_after"]) / abs(b["chi2_after"]) for a, b in zip(got.trace, ref.trace)] + [0.0])
        trials_equal = same_len and all(a["trials"] == b["trials"] and a["status"] == b["status"] for a, b in zip(got.trace, ref.trace))
        dpose = float(np.abs(got.pose_qt - ref.pose_qt).max())
        dpts = float(np.abs(got.points - ref.points[lo:hi]).max())
        levels_equal = bool(np.array_equal(got.edge_level, ref.edge_level[sel]))
        out = {"ok": bool(trials_equal and rel <= 1e-6 and dpose < 1e-6 and dpts < 1e-6 and levels_equal),
               "max_rel_chi2": rel, "max_abs_pose": dpose, "max_abs_point_rank0": dpts, "trials_equal": bool(trials_equal),
               "edge_levels_equal": levels_equal, "lm_iterations": len(got.trace),
               "map": "C4 at 1/10 size (%d observations), merge schedule, sharded over %d ranks, vs oracle/ba_ref on rank 0" % (full.n_obs, world)}
    return out


def measure(ctx, args, which: str, rank: int, world: int, l2_flush, steps: int, warmup: int, local_rank: int, clocks: bool):
    """Device-resident and end-to-end timing of one workload; returns the pieces of the JSON line (rank-reduced)."""
    from orb_slam3_study_kr_b200.problem import schedule_global_ba
    p, n_obs_total = make_workload(which, rank, world)
    ctx.pin_problem(p)
    s = schedule_global_ba(N_ITER)
    # ---- device-resident arm
    ctx.upload(p)
    for _ in range(warmup):
        ctx.reset_resident()
        ctx.solve_resident(s, download=False)
    barrier()
    dev_ms = 0.0
    acc = dict(build_ms=0.0, linsolve_ms=0.0, update_ms=0.0, build_launches=0, linsolve_launches=0, update_launches=0,
               total_launches=0, lm_iterations=0, lm_trials=0, edge_linearisations=0, edge_evaluations=0, solve_retries=0)
    t_wall0 = time.perf_counter()
    clk = ClockSampler(local_rank) if clocks else None
    if clk:
        clk.__enter__()
    for _ in range(steps):
        ctx.reset_resident()
        l2_flush()
        ctx.solve_resident(s, download=False)
        t = ctx.timing()
        dev_ms += t["solve_ms"]
        for k in acc:
            acc[k] += t[k]
    barrier()
    if clk:
        clk.__exit__()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    dev_ms_max = max_over_ranks(dev_ms)
    passes_all = sum_over_ranks(acc["edge_linearisations"] + acc["edge_evaluations"])
    timing_last = ctx.timing()

    # ---- end-to-end arm through bagpu_solve_ba with host buffers
    res_buf = ctx.alloc_result(p, s)                         # page-locked, reused: what a SLAM thread's adapter keeps
    for _ in range(min(warmup, 2)):
        ctx.solve_ba(p, s, into=res_buf)
    barrier()
    e2e_s, e2e_passes, h2d, d2h, e2e_phase = 0.0, 0, 0, 0, {}
    for _ in range(steps):
        l2_flush()
        barrier()
        t0 = time.perf_counter()
        ctx.solve_ba(p, s, into=res_buf)
        e2e_s += time.perf_counter() - t0
        t = ctx.timing()
        e2e_passes += t["edge_linearisations"] + t["edge_evaluations"]
        h2d, d2h = t["h2d_bytes"], t["d2h_bytes"]
        e2e_phase = {"upload_ms": t["h2d_ms"], "solve_ms": t["solve_ms"], "download_ms": t["d2h_ms"]}   # device-side phases of the last call
    e2e_value = sum_over_ranks(e2e_passes) / max_over_ranks(e2e_s)
    trace = res_buf[0].trace
    return dict(p=p, n_obs_total=n_obs_total, acc=acc, dev_ms=dev_ms, dev_ms_max=dev_ms_max, wall_ms=wall_ms, passes_all=passes_all,
                value=passes_all / (dev_ms_max * 1e-3), e2e_value=e2e_value, e2e_s=e2e_s, h2d=sum_over_ranks(h2d), d2h=sum_over_ranks(d2h),
                e2e_phase=e2e_phase, clocks=clk.summary() if clk else None, timing_last=timing_last, trace=trace)


def run_gpu(args, rank: int, world: int, local_rank: int):
    import torch
    from orb_slam3_study_kr_b200 import api

    torch.cuda.set_device(local_rank)
    ctx = api.Context(local_rank)
    if world > 1:
        uid = ctx.comm_unique_id() if rank == 0 else None
        uid = broadcast_bytes(uid, 128)
        ctx.comm_init(world, rank, uid)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def l2_flush():
        flush.fill_(1)
        torch.cuda.synchronize()

    parity = None if args.no_parity else parity_check(ctx, rank, world)
    which = args.workload
    m = measure(ctx, args, which, rank, world, l2_flush, args.steps, args.warmup, local_rank, clocks=True)
    p, acc, steps = m["p"], m["acc"], args.steps

    # ---- roofline of the linearise+Schur pass (this rank's share of the map)
    peak, peak_src = peaks()
    alg_bytes, alg_flops = pass_algorithmic(p, p.n_free, m["timing_last"]["schur_blocks"])
    build_avg_ms = acc["build_ms"] / max(1, acc["build_launches"])
    achieved = alg_bytes / (build_avg_ms * 1e-3) / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "build_pass_traffic.json")
    tensor_pct = None
    if os.path.exists(tf) and world == 1:
        try:
            prof = json.load(open(tf)).get(which, {})                                     # ncu --set full capture of the same workload
            traffic = prof.get("dram_bytes_per_launch")
            tensor_pct = prof.get("pair_tile_mma_kernel", {}).get("dmma_pipe_active_pct")
        except Exception:
            traffic = None
    fp64 = ctx.fp64_peak()

    also = None
    if not args.no_also:
        other = "c4" if which == "c5" else "c5"
        o = measure(ctx, args, other, rank, world, l2_flush, max(2, min(args.steps, 3)) if other == "c5" else args.steps, min(args.warmup, 3), local_rank, clocks=False)
        osteps = max(2, min(args.steps, 3)) if other == "c5" else args.steps
        also = {"workload": workload_name(other, o["n_obs_total"]), "value": o["value"], "unit": "edge passes/s", "steps": osteps,
                "ms_per_step": o["dev_ms_max"] / osteps, "lm_iters_per_s": o["acc"]["lm_iterations"] / (o["dev_ms_max"] * 1e-3),
                "lm_trials_per_step": o["acc"]["lm_trials"] / osteps,
                "e2e": {"value": o["e2e_value"], "ms_per_step": 1e3 * o["e2e_s"] / osteps, "phases_last_call": o["e2e_phase"]},
                "kernels_ms_per_step": {k: o["acc"][k + "_ms"] / osteps for k in ("build", "linsolve", "update")},
                "solver_parts": o["timing_last"]["solver_parts"]}

    if rank == 0:
        dev_ms = m["dev_ms"]
        line = {"metric": METRIC, "value": m["value"], "unit": "edge passes/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
                "ms_per_step": m["dev_ms_max"] / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(which, m["n_obs_total"]),
                           "sharding": ("contiguous landmark range per rank (%d observations on rank 0), NCCL all-reduce of the reduced camera system per LM trial" % p.n_obs) if world > 1 else "1 GPU",
                           "l2": "flushed between steps (256 MiB write)", "timing": "CUDA events on the library stream, max over ranks"},
                "lm_iters_per_s": acc["lm_iterations"] / (m["dev_ms_max"] * 1e-3), "lm_iterations_per_step": acc["lm_iterations"] / steps,
                "lm_trials_per_step": acc["lm_trials"] / steps, "wall_ms_per_step": m["wall_ms"] / steps,
                "ms_per_trial": m["dev_ms_max"] / max(1, acc["lm_trials"]),
                "e2e": {"value": m["e2e_value"], "unit": "edge passes/s", "h2d_bytes_per_step": int(m["h2d"]), "d2h_bytes_per_step": int(m["d2h"]),
                        "ms_per_step": 1e3 * m["e2e_s"] / steps, "phases_last_call": m["e2e_phase"]},
                "gpu_launches": int(acc["total_launches"]),
                "clocks": m["clocks"],
                "parity_check": parity,
                "roofline": {"kernel": "linearise+Schur pass: stage_kernel + tile_diag_kernel + pair_tile_mma_kernel (CUDA events around the launches of a pass on the library stream)", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": build_avg_ms,
                             "launches": int(acc["build_launches"]),
                             "note": "bytes count the 6x6 blocks inside the stored band (%d), not the dense pattern" % m["timing_last"]["schur_blocks"]},
                "roofline_fp64": {"kernel": "the same pass against the FP64 pipe (SURVEY 8d: max(bytes/BW, flops/peak))", "bound": "fp64",
                                  "achieved": alg_flops / (build_avg_ms * 1e-3) / 1e12, "peak": fp64["dfma_tflops"], "unit": "TFLOP/s",
                                  "frac": alg_flops / (build_avg_ms * 1e-3) / 1e12 / max(fp64["dfma_tflops"], 1e-9),
                                  "algorithmic_flops_per_launch": alg_flops, "peak_source": "measured on this device by bagpu_test_fp64_peak (DFMA probe)",
                                  "dmma_peak_tflops": fp64["dmma_tflops"],
                                  "tensor_pipe_active_pct_ncu": tensor_pct,      # pair_tile_mma_kernel, FP64 tensor pipe (profiles/r02g_full_raw_*.csv); not measured live
                                  "bound_time_ms": {"hbm": alg_bytes / (peak * 1e9) * 1e3, "fp64": alg_flops / (fp64["dfma_tflops"] * 1e12) * 1e3}},
                "kernels": {"build_ms_per_step": acc["build_ms"] / steps, "linsolve_ms_per_step": acc["linsolve_ms"] / steps,
                            "update_ms_per_step": acc["update_ms"] / steps,
                            "share_of_step": {k: acc[k + "_ms"] / dev_ms for k in ("build", "linsolve", "update")},
                            "solver_parts": m["timing_last"]["solver_parts"], "solve_retries": int(acc["solve_retries"])},
                "per_iteration": [{k: t[k] for k in ("iteration", "trials", "chi2_after", "active_edges", "linearise_schur_us", "linear_solve_us", "update_us", "iteration_us")}
                                  for t in m["trace"]],
                "also": also}
        if world == 1 and not args.no_local:
            line["local_ba"] = local_ba_lines(ctx, l2_flush)
        # CPU baseline beside it (rank 0, N=1 only): a bounded sample of the same map on one host core
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ba_ref
            from orb_slam3_study_kr_b200.problem import schedule_global_ba
            ref_iters = args.ref_iters if args.ref_iters > 0 else (2 if which == "c5" else 4)
            _, c = ba_ref.solve(p, schedule_global_ba(ref_iters), True)
            v = (c["edge_linearisations"] + c["edge_evaluations"]) / c["seconds"]
            line["cpu_baseline"] = {"value": v, "unit": "edge passes/s", "cores": 1, "kind": "port",
                                    "lm_iters_per_s": c["lm_iterations"] / c["seconds"],
                                    "sample": "%d LM iterations of the same %s map (%.1f s), oracle/ba_ref.cpp, 1 thread; host has %d cores"
                                              % (ref_iters, which.upper(), c["seconds"], os.cpu_count())}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="bagpu", choices=["bagpu", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-iters", type=int, default=0, help="LM iterations per CPU sample (0: 4 for c4; 1-2 for c5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-local", action="store_true", help="skip the secondary local-BA / pose-batch lines")
    ap.add_argument("--no-also", action="store_true", help="skip the other global-BA workload")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the (sharded) path before the timed region")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_gpu(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
