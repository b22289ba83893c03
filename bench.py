#!/usr/bin/env python
"""bench.py -- global bundle adjustment (BASELINE.json config 4) on N B200s, one JSON line.

  python bench.py --gpus N --steps K --warmup W            libbagpu arm
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle restatement, 1 thread)

Workload (config.workload): "C4 global BA": 500 keyframes (1 fixed), 200 000 landmarks and ~2.16 M stereo/mono
observations PER GPU (weak scaling: the keyframes are common, every rank owns its own 200 k landmarks and all their
observations), EuRoC intrinsics, seed 4, non-robust (LoopClosing's GBA call, src/LoopClosing.cc:2289), 20 LM iterations.
A step = one Optimizer::GlobalBundleAdjustemnt call = optimize(20) from the initial estimates.

  value   edge passes per second (edge linearisations + edge evaluations, summed over ranks) with the map resident in
          HBM, timed with CUDA events on the library's stream, max over ranks.
  e2e     the same metric through the reference-facing C-ABI call bagpu_solve_ba with HOST buffers: H2D of the whole
          problem, solve, D2H of poses / points / per-edge chi2 / flags inside the timed region.
  roofline  the linearise+Schur pass (stage_kernel + pair_kernel), the dominant kernel group of the library stream:
          algorithmic bytes per pass / mean pass duration (CUDA events around every pass in the timed region, on the
          library's stream) against the measured HBM copy bandwidth. On one GPU the band Cholesky runs BESIDE pair_kernel
          on a second stream, so the `kernels` shares overlap and sum to more than 1.
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


def workload_name(n_obs: int) -> str:
    return ("C4 global BA, weak-scaled: 500 KFs (1 fixed) common, 200k landmarks / %d observations per GPU, "
            "seed 4, EuRoC stereo+mono, non-robust, optimize(%d)" % (n_obs, N_ITER))


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
def make_workload(rank: int, world: int):
    from orb_slam3_study_kr_b200 import synthetic
    return synthetic.global_ba_shard(rank, world, robust=False)


def algorithmic_bytes_per_build_launch(p, n_free: int) -> float:
    """per observation 40 B (mono) / 48 B (stereo) read; per landmark 24 B read; the reduced camera
    system written once: 288 B per upper 6x6 block + 48 B per camera (dense pattern: all Nc(Nc+1)/2 blocks)."""
    from orb_slam3_study_kr_b200.problem import EDGE_STEREO
    n_st = int((p.obs_kind == EDGE_STEREO).sum())
    return 48.0 * n_st + 40.0 * (p.n_obs - n_st) + 24.0 * p.n_points + 288.0 * (n_free * (n_free + 1) / 2) + 48.0 * n_free


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
    bounded sample: `ref_iters` LM iterations of the same map."""
    if rank != 0:
        return
    from orb_slam3_study_kr_b200.problem import schedule_global_ba
    from oracle import ba_ref
    p = make_workload(0, 1)
    s = schedule_global_ba(args.ref_iters)
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
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "lm_iters_per_s": iters / tot_s,
            "config": {"workload": workload_name(p.n_obs),
                       "sample": "%d of %d LM iterations per step, rank 0's shard (the rate does not depend on N)" % (args.ref_iters, N_ITER)},
            "cpu_baseline": {"value": v, "unit": "edge passes/s", "cores": 1, "kind": "port",
                             "sample": "%d steps x %d LM iterations of the C4 map, oracle/ba_ref.cpp -O3 -march=x86-64-v3, 1 thread "
                                       "(reference g2o is built without OpenMP)" % (args.steps, args.ref_iters)},
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
def run_gpu(args, rank: int, world: int, local_rank: int):
    import torch
    from orb_slam3_study_kr_b200 import api
    from orb_slam3_study_kr_b200.problem import BAResult, schedule_global_ba

    torch.cuda.set_device(local_rank)
    ctx = api.Context(local_rank)
    if world > 1:
        uid = ctx.comm_unique_id() if rank == 0 else None
        uid = broadcast_bytes(uid, 128)
        ctx.comm_init(world, rank, uid)
    p = make_workload(rank, world)
    ctx.pin_problem(p)
    s = schedule_global_ba(N_ITER)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def l2_flush():
        flush.fill_(1)
        torch.cuda.synchronize()

    # ---- device-resident arm
    ctx.upload(p)
    for _ in range(args.warmup):
        ctx.reset_resident()
        ctx.solve_resident(s, download=False)
    barrier()
    dev_ms = 0.0
    acc = dict(build_ms=0.0, linsolve_ms=0.0, update_ms=0.0, build_launches=0, linsolve_launches=0, update_launches=0,
               total_launches=0, lm_iterations=0, lm_trials=0, edge_linearisations=0, edge_evaluations=0)
    t_wall0 = time.perf_counter()
    with ClockSampler(local_rank) as clk:
        for _ in range(args.steps):
            ctx.reset_resident()
            l2_flush()
            ctx.solve_resident(s, download=False)
            t = ctx.timing()
            dev_ms += t["solve_ms"]
            for k in acc:
                acc[k] += t[k]
        barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    dev_ms_max = max_over_ranks(dev_ms)
    passes_all = sum_over_ranks(acc["edge_linearisations"] + acc["edge_evaluations"])
    value = passes_all / (dev_ms_max * 1e-3)
    iters_per_s = acc["lm_iterations"] / (dev_ms_max * 1e-3)

    # ---- end-to-end arm through bagpu_solve_ba with host buffers
    res_buf = ctx.alloc_result(p, s)                         # page-locked, reused: what a SLAM thread's adapter keeps
    for _ in range(min(args.warmup, 2)):
        ctx.solve_ba(p, s, into=res_buf)
    barrier()
    e2e_s = 0.0
    e2e_passes = 0
    h2d = d2h = 0
    for _ in range(args.steps):
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

    # ---- roofline of the linearise+Schur kernel
    peak, peak_src = peaks()
    alg_bytes = algorithmic_bytes_per_build_launch(p, p.n_free)
    build_avg_ms = acc["build_ms"] / max(1, acc["build_launches"])
    achieved = alg_bytes / (build_avg_ms * 1e-3) / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "build_pass_traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "edge passes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(p.n_obs),
                           "sharding": "landmarks per rank, NCCL all-reduce of the reduced camera system per LM trial" if world > 1 else "1 GPU",
                           "l2": "flushed between steps (256 MiB write)", "timing": "CUDA events on the library stream, max over ranks"},
                "lm_iters_per_s": iters_per_s, "lm_iterations_per_step": acc["lm_iterations"] / args.steps,
                "lm_trials_per_step": acc["lm_trials"] / args.steps, "wall_ms_per_step": wall_ms / args.steps,
                "e2e": {"value": e2e_value, "unit": "edge passes/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": 1e3 * e2e_s / args.steps, "phases_last_call": e2e_phase},
                "gpu_launches": int(acc["total_launches"]),
                "clocks": clk.summary(),
                "roofline": {"kernel": "linearise+Schur pass: stage_kernel + pair_kernel (CUDA events around the two launches on the library stream; the band Cholesky runs beside pair_kernel on its own stream)", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": build_avg_ms,
                             "launches": int(acc["build_launches"])},
                "kernels": {"build_ms_per_step": acc["build_ms"] / args.steps, "linsolve_ms_per_step": acc["linsolve_ms"] / args.steps,
                            "update_ms_per_step": acc["update_ms"] / args.steps,
                            "share_of_step": {k: acc[k + "_ms"] / dev_ms for k in ("build", "linsolve", "update")}}}
        if world == 1 and not args.no_local:
            line["local_ba"] = local_ba_lines(ctx, l2_flush)
        # CPU baseline beside it (rank 0, N=1 only): a bounded sample of the same map on one host core
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ba_ref
            _, c = ba_ref.solve(p, schedule_global_ba(args.ref_iters), True)
            v = (c["edge_linearisations"] + c["edge_evaluations"]) / c["seconds"]
            line["cpu_baseline"] = {"value": v, "unit": "edge passes/s", "cores": 1, "kind": "port",
                                    "lm_iters_per_s": c["lm_iterations"] / c["seconds"],
                                    "sample": "%d LM iterations of the same C4 map (%.1f s), oracle/ba_ref.cpp, 1 thread; host has %d cores"
                                              % (args.ref_iters, c["seconds"], os.cpu_count())}
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
    ap.add_argument("--ref-iters", type=int, default=4, help="LM iterations per CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-local", action="store_true", help="skip the secondary local-BA / pose-batch lines")
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
