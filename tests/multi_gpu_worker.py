"""torchrun worker: landmark-sharded global BA over NCCL vs the same map solved on one GPU (and vs the oracle).
Launched by tests/test_gpu_multi.py or by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_worker.py"""
import faulthandler
import os
import sys

faulthandler.dump_traceback_later(int(os.environ.get("WORKER_WATCHDOG_S", "240")), exit=True)   # a hang prints where it stands and exits non-zero

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from orb_slam3_study_kr_b200 import api, problem, synthetic  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = api.Context(local)
uid = ctx.comm_unique_id() if rank == 0 else None
ctx.comm_init(world, rank, bench.broadcast_bytes(uid, 128))

n_kf, ppr = 60, 3000
shards = [synthetic.global_ba_shard(r, world, n_kf=n_kf, points_per_rank=ppr, robust=(r % 2 == 0)) for r in range(world)]
mine = shards[rank]
s = problem.schedule_merge_ba()            # two rounds + gate: exercises every collective
got = ctx.solve_ba(mine, s)
print(f"[rank {rank}] sharded solve done: {len(got.trace)} iterations", flush=True)   # progress markers: a time-out shows where it stopped

# rank 0 compares; a failure is reported AFTER the collective teardown below, so that no rank is left waiting for another
failure = None
if rank == 0:
    try:
        full = synthetic.concat_shards(shards)
        solo = api.Context(local)
        ref = solo.solve_ba(full, s)
        print("[rank 0] single-GPU solve of the concatenated map done", flush=True)
        from oracle import ba_ref
        orc = ba_ref.solve(full, s)
        print("[rank 0] oracle done", flush=True)
        assert len(got.trace) == len(ref.trace) == len(orc.trace), (len(got.trace), len(ref.trace), len(orc.trace))
        for a, b, c in zip(got.trace, ref.trace, orc.trace):
            assert a["trials"] == b["trials"] == c["trials"] and a["status"] == c["status"]
            assert abs(a["chi2_after"] - c["chi2_after"]) <= 1e-6 * c["chi2_after"], (a, c)          # the bar: the oracle
            # one GPU vs two: the same sums in a different order (partial systems are added by NCCL), 1e-8 late in the second round
            assert abs(a["chi2_after"] - b["chi2_after"]) <= 1e-8 * b["chi2_after"], (a, b)
        assert np.abs(got.pose_qt - orc.pose_qt).max() < 1e-6
        lo, hi = 0, mine.n_points
        assert np.abs(got.points - orc.points[lo:hi]).max() < 1e-6
        assert np.array_equal(got.edge_level, orc.edge_level[: mine.n_obs])
        print("multi-gpu ok: world", world, "iterations", len(got.trace), "final chi2", got.trace[-1]["chi2_after"], flush=True)
        solo.close()
    except Exception as e:              # noqa: BLE001
        import traceback
        failure = traceback.format_exc()
# every rank holds the same poses afterwards
t = torch.from_numpy(got.pose_qt.copy()).cuda()
t0 = t.clone(); dist.broadcast(t0, 0)
assert torch.equal(t, t0), "poses differ across ranks"
dist.barrier()
print(f"[rank {rank}] poses agree across ranks", flush=True)
ctx.close()
dist.destroy_process_group()
if failure:
    print(failure, flush=True)
    sys.exit(1)
