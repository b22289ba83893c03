// Compiled C++ adapter test: fills bagpu_problem the way INTEGRATION.md's FlatBA / PushPose / PushEdge do (std::vector
// gather buffers, poses and points in ascending id, edges landmark by landmark), makes ONE bagpu_solve_ba call with the
// LocalMapping schedule, and runs the classification loop of src/Optimizer.cc:1416-1460 on the returned per-edge values.
// Built and run by tests/test_abi_and_host.py::test_cpp_adapter (g++ tests/abi/adapter.cpp -lbagpu). No Python, no torch.
//
// The scene is a small synthetic window (pinhole EuRoC intrinsics, mono + stereo edges, 5 % gross outliers); the program
// prints one JSON line the Python side checks: statuses, chi2 before/after, how many injected outliers the loop erases.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <utility>
#include <vector>

#include "bagpu.h"

namespace {
struct FlatBA {                                // what the g2o graph used to hold
    std::vector<double> pose_qt, points, u, v, ur, inv_sigma2;
    std::vector<uint8_t> pose_fixed, kind, flags;
    std::vector<int32_t> obs_pose, obs_point;
    std::vector<int16_t> obs_cam, obs_rig;
    std::vector<bagpu_camera> cams;
    std::vector<std::pair<int, int>> edge_owner;   // (keyframe, map point): what vpEdgeKF* / vpMapPointEdge* held
    std::vector<uint8_t> injected_outlier;
    bagpu_problem view() const {
        bagpu_problem p = {};
        p.n_poses = (int32_t)pose_fixed.size(); p.pose_qt = pose_qt.data(); p.pose_fixed = pose_fixed.data();
        p.n_points = (int32_t)(points.size() / 3); p.points = points.data();
        p.n_cameras = (int32_t)cams.size(); p.cameras = cams.data(); p.n_rigs = 0; p.rigs = nullptr;
        p.n_obs = (int64_t)u.size(); p.obs_pose = obs_pose.data(); p.obs_point = obs_point.data(); p.obs_cam = obs_cam.data();
        p.obs_rig = obs_rig.data(); p.obs_kind = kind.data(); p.obs_flags = flags.data(); p.obs_u = u.data(); p.obs_v = v.data();
        p.obs_ur = ur.data(); p.obs_inv_sigma2 = inv_sigma2.data();
        return p;
    }
};
}  // namespace

int main() {
    std::mt19937_64 rng(7);
    std::normal_distribution<double> N01(0.0, 1.0);
    std::uniform_real_distribution<double> U01(0.0, 1.0);
    const float fx = 458.654f, fy = 457.296f, cx = 367.215f, cy = 248.375f, bf = 458.654f * 0.110074f;
    FlatBA f;
    bagpu_camera cam = {};
    cam.type = BAGPU_CAM_PINHOLE; cam.p[0] = fx; cam.p[1] = fy; cam.p[2] = cx; cam.p[3] = cy; cam.bf = bf;
    f.cams.push_back(cam);
    // keyframes on a line along x, looking down +z (identity rotation): Tcw = [I | -c]; the first 4 are fixed
    const int n_kf = 16, n_fixed = 4, n_pt = 1500;
    std::vector<double> cxw(n_kf);
    for (int k = 0; k < n_kf; k++) {
        cxw[k] = 0.15 * k;
        const bool fixed = k < n_fixed;
        const double tx = -cxw[k] + (fixed ? 0.0 : 0.01 * N01(rng)), ty = fixed ? 0.0 : 0.01 * N01(rng), tz = fixed ? 0.0 : 0.01 * N01(rng);
        const double qx = fixed ? 0.0 : 0.004 * N01(rng), qy = fixed ? 0.0 : 0.004 * N01(rng), qz = fixed ? 0.0 : 0.004 * N01(rng);
        const double pose[7] = {(double)(float)tx, (double)(float)ty, (double)(float)tz, (double)(float)qx, (double)(float)qy, (double)(float)qz, 1.0};
        f.pose_qt.insert(f.pose_qt.end(), pose, pose + 7);     // not normalised: the library normalises like SE3Quat(q, t)
        f.pose_fixed.push_back(fixed);
    }
    const float sig2[4] = {1.0f, 0.6944444179534912f, 0.4822530746459961f, 0.33489790558815f};
    for (int j = 0; j < n_pt; j++) {
        const double X = -1.0 + 4.4 * U01(rng), Y = -1.5 + 3.0 * U01(rng), Z = 3.0 + 6.0 * U01(rng);
        const double P0[3] = {(double)(float)(X + 0.02 * N01(rng)), (double)(float)(Y + 0.02 * N01(rng)), (double)(float)(Z + 0.02 * N01(rng))};
        f.points.insert(f.points.end(), P0, P0 + 3);
        for (int k = 0; k < n_kf; k++) {                        // pMP->GetObservations(), keyframe ascending
            const double xc = X - cxw[k], yc = Y, zc = Z;
            const double uu = fx * xc / zc + cx, vv = fy * yc / zc + cy;
            if (uu < 10 || uu > 742 || vv < 10 || vv > 470 || U01(rng) < 0.35) continue;
            const int lvl = (int)(U01(rng) * 4) & 3;
            const double sd = std::pow(1.2, lvl);
            const bool out = U01(rng) < 0.05;
            const bool stereo = U01(rng) < 0.6;
            const double du = sd * N01(rng) + (out ? 15.0 + 15.0 * U01(rng) : 0.0), dv = sd * N01(rng) - (out ? 15.0 + 15.0 * U01(rng) : 0.0);
            f.obs_pose.push_back(k); f.obs_point.push_back(j); f.obs_cam.push_back(0); f.obs_rig.push_back(-1);
            f.kind.push_back(stereo ? BAGPU_EDGE_STEREO : BAGPU_EDGE_MONO); f.flags.push_back(BAGPU_FLAG_ROBUST);
            f.u.push_back((double)(float)(uu + du)); f.v.push_back((double)(float)(vv + dv));
            f.ur.push_back(stereo ? (double)(float)(uu + du - bf / zc + sd * N01(rng)) : 0.0);
            f.inv_sigma2.push_back((double)sig2[lvl]);
            f.edge_owner.emplace_back(k, j); f.injected_outlier.push_back(out);
        }
    }
    bagpu_ctx *ctx = nullptr;
    int rc = bagpu_init(-1, &ctx);
    if (rc != BAGPU_OK) { std::printf("{\"error\": \"bagpu_init: %s\"}\n", bagpu_strerror(rc)); return 2; }

    // ONE call: optimizer.initializeOptimization(); optimizer.optimize(10);   (Optimizer.cc:1410-1411)
    volatile uint8_t stop_flag = 0;
    const bagpu_round round = {10, BAGPU_GATE_NONE, 5.991, 7.815, 0, 0};
    const float thHuberMono = std::sqrt(5.991f), thHuberStereo = std::sqrt(7.815f);
    bagpu_trace trace[16];
    bagpu_schedule s = {1, &round, thHuberMono, thHuberStereo, 0.0, &stop_flag, BAGPU_SOLVER_AUTO, 16};
    std::vector<double> pose(f.pose_qt.size()), pts(f.points.size()), chi2(f.u.size());
    std::vector<uint8_t> depth(f.u.size()), level(f.u.size());
    bagpu_result r = {pose.data(), pts.data(), chi2.data(), depth.data(), level.data(), trace, 0, 0};
    const bagpu_problem p = f.view();
    rc = bagpu_solve_ba(ctx, &p, &s, &r);
    if (rc < 0) { std::printf("{\"error\": \"bagpu_solve_ba: %s: %s\"}\n", bagpu_strerror(rc), bagpu_last_error(ctx)); bagpu_destroy(ctx); return 3; }

    // classification exactly as Optimizer.cc:1416-1460, from the returned per-edge values
    std::vector<std::pair<int, int>> vToErase;
    long caught = 0, injected = 0, false_pos = 0;
    for (size_t e = 0; e < chi2.size(); ++e) {
        const double th = (f.kind[e] == BAGPU_EDGE_STEREO) ? 7.815 : 5.991;
        const bool erase = chi2[e] > th || !depth[e];
        if (erase) vToErase.push_back(f.edge_owner[e]);
        injected += f.injected_outlier[e];
        caught += erase && f.injected_outlier[e];
        false_pos += erase && !f.injected_outlier[e];
    }
    // a bad argument must come back as an error code with a message, never crash or throw
    bagpu_problem bad = p; bad.n_obs = 5; bad.obs_u = nullptr;
    const int rc_bad = bagpu_solve_ba(ctx, &bad, &s, &r);
    // the stop flag set before the call: nothing runs (Optimizer.cc:1406-1408)
    stop_flag = 1;
    bagpu_result r2 = {pose.data(), pts.data(), chi2.data(), depth.data(), level.data(), trace + 12, 0, 0};
    bagpu_schedule s2 = s; s2.max_trace = 4;
    const int rc_stop = bagpu_solve_ba(ctx, &p, &s2, &r2);
    std::printf("{\"rc\": %d, \"n_trace\": %d, \"chi2_first\": %.9e, \"chi2_last\": %.9e, \"edges\": %zu, \"injected\": %ld, \"caught\": %ld, "
                "\"false_pos\": %ld, \"erased\": %zu, \"rc_bad\": %d, \"rc_stop\": %d, \"stop_trace\": %d, \"active_edges\": %lld, \"version\": %d}\n",
                rc, r.n_trace, trace[0].chi2_before, trace[r.n_trace - 1].chi2_after, chi2.size(), injected, caught, false_pos,
                vToErase.size(), rc_bad, rc_stop, r2.n_trace, (long long)trace[0].active_edges, BAGPU_VERSION);
    bagpu_destroy(ctx);
    return 0;
}
