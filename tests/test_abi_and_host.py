"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every symbol include/bagpu.h
declares (no compute without a GPU), struct layouts match, the product path fails loudly without a device, the
generators hit the BASELINE shapes, and the landmark sharding used for multi-GPU global BA is a partition
(checked across 2 gloo ranks)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from orb_slam3_study_kr_b200 import api, problem, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bagpu.h")).read()
    declared = sorted(set(re.findall(r"\b(bagpu_[a-z0-9_]+)\s*\(", hdr)))
    assert set(declared) == set(api.EXPORTS), (declared, api.EXPORTS)
    lib = api.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bagpu_strerror(0).decode() == "ok"
    assert b"stop" in lib.bagpu_strerror(3)


def test_struct_layouts_match_header():
    """sizeof() of every ctypes mirror equals the C compiler's view of include/bagpu.h."""
    src = '#include "%s/include/bagpu.h"\n#include <stdio.h>\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n",' \
          'sizeof(bagpu_camera),sizeof(bagpu_rig),sizeof(bagpu_problem),sizeof(bagpu_round),sizeof(bagpu_schedule),' \
          'sizeof(bagpu_trace),sizeof(bagpu_result),sizeof(bagpu_pose_batch),sizeof(bagpu_pose_result),sizeof(bagpu_timing));return 0;}' % ROOT
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    mirrors = [problem.CCamera, problem.CRig, problem.CProblem, problem.CRound, problem.CSchedule, problem.CTrace,
               problem.CResult, problem.CPoseBatch, problem.CPoseResult, problem.CTiming]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of falling back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.BagpuError):
        api.Context(0)
    # the package itself never imports the oracle
    for f in os.listdir(os.path.join(ROOT, "orb_slam3_study_kr_b200")):
        if f.endswith(".py"):
            txt = open(os.path.join(ROOT, "orb_slam3_study_kr_b200", f)).read()
            assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", ""), f
    for f in os.listdir(os.path.join(ROOT, "orb_slam3_study_kr_b200", "csrc")):
        if not os.path.isfile(os.path.join(ROOT, "orb_slam3_study_kr_b200", "csrc", f)):
            continue
        assert "ba_ref" not in open(os.path.join(ROOT, "orb_slam3_study_kr_b200", "csrc", f), errors="ignore").read(), f


def test_generators_hit_baseline_shapes():
    p = synthetic.config(1)
    assert (p.n_poses, p.n_free, p.n_points) == (30, 20, 5000) and 36000 < p.n_obs < 46000
    assert (p.obs_kind == problem.EDGE_MONO).all() and p.cameras[0]["type"] == problem.CAM_PINHOLE
    p = synthetic.config(2)
    assert (p.n_poses, p.n_free, p.n_points) == (70, 45, 12000) and 105000 < p.n_obs < 135000
    fs = (p.obs_kind == problem.EDGE_STEREO).mean()
    assert 0.6 < fs < 0.75
    p = synthetic.config(3)
    assert (p.n_poses, p.n_free, p.n_points) == (45, 30, 8000) and 90000 < p.n_obs < 125000
    assert set(np.unique(p.obs_kind)) == {problem.EDGE_MONO, problem.EDGE_BODY}
    assert p.cameras[0]["type"] == problem.CAM_KB8 and p.rigs.shape == (1, 7)
    # both eyes of one keyframe can see the same point: two edges on one (pose, point) pair
    key = p.obs_point.astype(np.int64) * 1000 + p.obs_pose
    assert (np.diff(key) == 0).any()
    # landmark-major, pose-ascending order (what Optimizer.cc's per-MapPoint loops produce)
    assert (np.diff(key) >= 0).all()
    p = synthetic.config(4, scale=0.05)
    assert p.pose_fixed.sum() == 1 and p.pose_fixed[0] == 1
    # inputs live on the float grid, as at the reference's boundary
    for a in (p.pose_qt, p.points, p.obs_u, p.obs_v, p.obs_ur):
        assert np.array_equal(a, a.astype(np.float32).astype(np.float64))


def test_schedules_mirror_the_reference():
    s = problem.schedule_local_ba()
    assert [r.iterations for r in s.rounds] == [10] and s.delta_mono == problem.DELTA_MONO_LBA and s.lambda_init == 0
    assert problem.schedule_local_ba(inertial=True).lambda_init == 100.0
    s = problem.schedule_merge_ba()
    assert [r.iterations for r in s.rounds] == [5, 10] and s.rounds[0].gate_after == problem.GATE_LBA
    assert s.rounds[0].drop_kernel_after and s.delta_mono == problem.DELTA_MONO_GBA
    assert (s.rounds[0].gate_mono, s.rounds[0].gate_stereo) == (5.991, 7.815)
    s = problem.schedule_global_ba(20)
    assert [r.iterations for r in s.rounds] == [20] and s.rounds[0].gate_after == problem.GATE_NONE


def test_shard_by_landmark_is_a_partition():
    p = synthetic.config(4, scale=0.03)
    seen = np.zeros(p.n_obs, int)
    npts = 0
    for r in range(4):
        sh = p.shard_by_landmark(r, 4)
        seen += sh.truth["obs_mask"]
        npts += sh.n_points
        assert sh.n_poses == p.n_poses and (sh.obs_point >= 0).all() and (sh.obs_point < sh.n_points).all()
        lo, hi = sh.truth["point_range"]
        assert np.array_equal(sh.points, p.points[lo:hi])
    assert (seen == 1).all() and npts == p.n_points


_GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from orb_slam3_study_kr_b200 import synthetic, problem
import bench
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1], rank=int(sys.argv[2]), world_size=2)
rank = dist.get_rank()
p = synthetic.config(4, scale=0.03)
sh = p.shard_by_landmark(rank, 2)
# every observation lands on exactly one rank
t = torch.tensor([sh.n_obs, sh.n_points], dtype=torch.int64); dist.all_reduce(t)
assert t.tolist() == [p.n_obs, p.n_points], t
# the 128-byte communicator id travels from rank 0 exactly as bench.py moves it
uid = bytes(range(128)) if rank == 0 else None
got = bench.broadcast_bytes(uid, 128)
assert got == bytes(range(128))
# max-over-ranks timing helper
assert bench.max_over_ranks(1.0 + rank) == 2.0
# a camera-system style reduction: per-rank partial sums over the shard add up to the global sum
part = np.zeros(6 * p.n_poses); np.add.at(part, 6 * sh.obs_pose, sh.obs_u)
tt = torch.from_numpy(part); dist.all_reduce(tt)
full = np.zeros(6 * p.n_poses); np.add.at(full, 6 * p.obs_pose, p.obs_u)
assert np.allclose(tt.numpy(), full, rtol=1e-12)
dist.destroy_process_group()
print("ok", rank)
'''


def test_world_size_2_gloo_sharding(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % ROOT)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [pr.communicate(timeout=240)[0].decode() for pr in procs]
    for pr, o in zip(procs, outs):
        assert pr.returncode == 0, o


def build_cpp_adapter(tmp_path):
    """g++ tests/abi/adapter.cpp -lbagpu: a compiled consumer of include/bagpu.h, filled the way INTEGRATION.md does."""
    exe = str(tmp_path / "adapter")
    libdir = os.path.join(ROOT, "orb_slam3_study_kr_b200")
    api.load_library()                                        # fails loudly if the library has not been built
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi", "adapter.cpp"), "-o", exe, "-L" + libdir, "-lbagpu",
                           "-Wl,-rpath," + libdir])
    return exe


def test_cpp_adapter_compiles_and_refuses_without_gpu(tmp_path):
    """The C++ adapter links against the C ABI (no C++ types cross it); without a device bagpu_init reports
    BAGPU_ERR_NO_DEVICE and nothing is computed."""
    import json
    import torch
    exe = build_cpp_adapter(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_boundary.py")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 2 and "no CUDA device" in json.loads(out.stdout)["error"]


def test_strong_scaled_bench_workload_is_a_partition_of_the_config():
    """bench.py's N-rank workload: the union of the ranks' landmark ranges is exactly the full map (checked on config 4 at
    1/20 size through the same code path)."""
    p = synthetic.config(4, scale=0.05)
    tot_obs = tot_pts = 0
    for r in range(8):
        sh = p.shard_by_landmark(r, 8)
        tot_obs += sh.n_obs
        tot_pts += sh.n_points
        # MapPoint ids grow with the keyframe that created them: a contiguous landmark range touches a contiguous keyframe range
        kfs = np.unique(sh.obs_pose)
        assert kfs.max() - kfs.min() < p.n_poses // 8 + 2 * 40 + 2
    assert (tot_obs, tot_pts) == (p.n_obs, p.n_points)
