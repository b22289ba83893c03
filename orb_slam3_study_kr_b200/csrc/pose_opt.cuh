// Batched Optimizer::PoseOptimization (src/Optimizer.cc:815-1114): one CTA per frame, lanes stride over the
// frame's unary edges, the whole 4-round x optimize(10) Levenberg-Marquardt loop runs on the device with
// no host round trip. Per round: reset to the initial pose (:1008-1009), initializeOptimization(0), LM with the
// 6x6 dense solve (LinearSolverDense), then the float chi2 gate with recomputation of current outliers
// (:1013-1100); the Huber kernel is dropped after round index 2 (:1040-1041).
#pragma once
#include <float.h>
#include "ba_math.cuh"

#ifndef PO_THREADS
#define PO_THREADS 128
#endif
#ifndef PO_MINB
#define PO_MINB 1
#endif
#define PO_WARPS (PO_THREADS / 32)

struct PoseDev {
    int n_frames;
    const double *pose0;          // [F][7] normalised
    const int64_t *frame_ptr;     // [F+1]
    const bagpu_camera *cams;
    const double *rigs;
    const double *xw;             // [Ne][3]
    const uint32_t *o_meta;       // kind | cam<<2 | rig<<10
    const double *o_u, *o_v, *o_ur, *o_w;
    double delta_mono, delta_stereo;
    float gate_mono, gate_stereo;
    double *chi2;                 // [Ne] scratch: e->chi2() as last computed
    uint8_t *outlier;             // [Ne] out; doubles as the edge level
    double *pose_out;             // [F][7]
    int *n_inliers;               // [F]
    double *final_chi2;           // [F]
};

// block-wide sum of NV doubles per thread; result valid in every thread (fixed order: lanes by butterfly, warps 0..3)
template <int NV>
BA_DEV void block_allsum(double *v, double (*sh)[PO_WARPS]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = warp_allsum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) sh[i][warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < PO_WARPS; w++) s += sh[i][w];
        v[i] = s;
    }
}

// LDL^T solve of the 6x6 SPD system (H + lambda I) x = b; false if a pivot is not positive
// (LinearSolverDense: Eigen::LDLT + isPositive(), linear_solver_dense.h:101-107).
BA_DEV bool solve6(const double *H21 /*upper, row-major packed 21*/, double lambda, const double *b, double *x) {
    double A[6][6];
    int p = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = r; c < 6; c++) { A[r][c] = H21[p]; A[c][r] = H21[p]; p++; }
#pragma unroll
    for (int r = 0; r < 6; r++) A[r][r] += lambda;
    double d[6];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        double dk = A[k][k];
#pragma unroll
        for (int m = 0; m < k; m++) dk -= A[k][m] * A[k][m] * d[m];
        d[k] = dk;
        if (!(dk > 0.0)) ok = false;
#pragma unroll
        for (int i = k + 1; i < 6; i++) {
            double s = A[i][k];
#pragma unroll
            for (int m = 0; m < k; m++) s -= A[i][m] * A[k][m] * d[m];
            A[i][k] = s / dk;
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { double s = b[i];
#pragma unroll
        for (int m = 0; m < i; m++) s -= A[i][m] * y[m]; y[i] = s; }
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] /= d[i];
#pragma unroll
    for (int i = 5; i >= 0; i--) { double s = y[i];
#pragma unroll
        for (int m = i + 1; m < 6; m++) s -= A[m][i] * x[m]; x[i] = s; }
    return ok;
}

__global__ void __launch_bounds__(PO_THREADS, PO_MINB) pose_opt_kernel(PoseDev D) {
    __shared__ double sh[28][PO_WARPS];
    __shared__ double s_T[7];
    __shared__ int s_flag;
    const int f = blockIdx.x;
    if (f >= D.n_frames) return;
    const int64_t e0 = D.frame_ptr[f], e1 = D.frame_ptr[f + 1];
    const int n = (int)(e1 - e0);
    const Pose T0 = load_pose(D.pose0 + 7 * (size_t)f);
    for (int i = threadIdx.x; i < n; i += PO_THREADS) D.outlier[e0 + i] = 0;
    if (n < 3) {                                                      // Optimizer.cc:996-997
        if (threadIdx.x == 0) { store_pose(D.pose_out + 7 * (size_t)f, T0); D.n_inliers[f] = 0; D.final_chi2[f] = 0.0; }
        return;
    }
    __syncthreads();
    Pose T = T0;
    bool robust = true;
    int nBadGate = 0;
    double lastChi = 0.0;

    // evaluate edge i at pose P: residual, chi2
    auto eval = [&](int i, const Pose &P, double &r0, double &r1, double &r2, double &om, int &kind, Cam &cam, Pose &Trl) {
        const int64_t e = e0 + i;
        const uint32_t m = D.o_meta[e];
        kind = META_KIND(m);
        cam = load_cam(D.cams + META_CAM(m));
        if (kind == BAGPU_EDGE_BODY) Trl = load_pose(D.rigs + 7 * META_RIG(m));
        om = D.o_w[e];
        edge_residual(kind, cam, P, &Trl, D.xw[3 * e], D.xw[3 * e + 1], D.xw[3 * e + 2], D.o_u[e], D.o_v[e],
                      (kind == BAGPU_EDGE_STEREO) ? D.o_ur[e] : 0.0, true, r0, r1, r2);
    };

    for (int it = 0; it < 4; it++) {
        T = T0;
        double cnt[1] = {0.0};
        for (int i = threadIdx.x; i < n; i += PO_THREADS) cnt[0] += (D.outlier[e0 + i] == 0) ? 1.0 : 0.0;
        block_allsum<1>(cnt, sh);
        if (cnt[0] > 0.5) {
            double lambda = -1.0, ni = 2.0;
            int nb = 0;
            bool ok = true;
            for (int k = 0; k < 10 && ok; k++) {
                // computeActiveErrors + buildSystem at T
                double acc[28];
#pragma unroll
                for (int q = 0; q < 28; q++) acc[q] = 0.0;
                for (int i = threadIdx.x; i < n; i += PO_THREADS) {
                    if (D.outlier[e0 + i]) continue;
                    double r0, r1, r2, om; int kind; Cam cam; Pose Trl;
                    eval(i, T, r0, r1, r2, om, kind, cam, Trl);
                    const double chi2 = r0 * (om * r0) + r1 * (om * r1) + r2 * (om * r2);
                    D.chi2[e0 + i] = chi2;
                    double rho0 = chi2, rho1 = 1.0;
                    if (robust) huber(chi2, kind == BAGPU_EDGE_STEREO ? D.delta_stereo : D.delta_mono, rho0, rho1);
                    EdgeLin L;
                    const int64_t e = e0 + i;
                    edge_linearize(kind, cam, T, &Trl, D.xw[3 * e], D.xw[3 * e + 1], D.xw[3 * e + 2], L);
                    const double w = rho1 * om;
                    const double *B = L.B;
                    int p = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
#pragma unroll
                        for (int c = a; c < 6; c++) { acc[p] += w * (B[a] * B[c] + B[6 + a] * B[6 + c] + B[12 + a] * B[12 + c]); p++; }
                        acc[21 + a] -= rho1 * (B[a] * (om * r0) + B[6 + a] * (om * r1) + B[12 + a] * (om * r2));
                    }
                    acc[27] += rho0;
                }
                block_allsum<28>(acc, sh);
                double currentChi = acc[27], tempChi = currentChi;
                const double iniChi = currentChi;
                if (k == 0) {
                    const double m = fmax(fmax(fabs(acc[0]), fabs(acc[6])), fmax(fmax(fabs(acc[11]), fabs(acc[15])), fmax(fabs(acc[18]), fabs(acc[20]))));
                    lambda = 1e-5 * m; ni = 2.0; nb = 0;
                }
                double rho = 0.0;
                int qmax = 0;
                do {
                    // every thread solves the same 6x6 system (identical inputs -> identical x), no broadcast needed
                    double x[6] = {0, 0, 0, 0, 0, 0};
                    const bool ok2 = solve6(acc, lambda, acc + 21, x);
                    if (!ok2) { for (int q = 0; q < 6; q++) x[q] = 0.0; }
                    const Pose Tt = pose_oplus(T, x);
                    double tc[1] = {0.0};
                    for (int i = threadIdx.x; i < n; i += PO_THREADS) {
                        if (D.outlier[e0 + i]) continue;
                        double r0, r1, r2, om; int kind; Cam cam; Pose Trl;
                        eval(i, Tt, r0, r1, r2, om, kind, cam, Trl);
                        const double chi2 = r0 * (om * r0) + r1 * (om * r1) + r2 * (om * r2);
                        D.chi2[e0 + i] = chi2;
                        double rho0 = chi2, rho1;
                        if (robust) huber(chi2, kind == BAGPU_EDGE_STEREO ? D.delta_stereo : D.delta_mono, rho0, rho1);
                        tc[0] += rho0;
                    }
                    block_allsum<1>(tc, sh);
                    tempChi = ok2 ? tc[0] : DBL_MAX;
                    rho = currentChi - tempChi;
                    double scale = 0.0;
#pragma unroll
                    for (int q = 0; q < 6; q++) scale += x[q] * (lambda * x[q] + acc[21 + q]);
                    scale += 1e-3;
                    rho /= scale;
                    if (rho > 0 && isfinite(tempChi)) {
                        const double t = 2 * rho - 1;
                        double alpha = 1. - t * t * t;
                        alpha = fmin(alpha, 2. / 3.);
                        lambda *= fmax(1. / 3., alpha);
                        ni = 2.0;
                        currentChi = tempChi;
                        T = Tt;
                    } else {
                        lambda *= ni;
                        ni *= 2.0;
                    }
                    qmax++;
                } while (rho < 0 && qmax < 10);
                lastChi = currentChi;
                if (qmax == 10 || rho == 0) ok = false;
                else {
                    if ((iniChi - currentChi) * 1e3 < iniChi) nb++; else nb = 0;
                    if (nb >= 3) ok = false;
                }
            }
        }
        // gate
        double bad[1] = {0.0};
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += PO_THREADS) {
            const int64_t e = e0 + i;
            double c = D.chi2[e];
            const int kind = META_KIND(D.o_meta[e]);
            if (D.outlier[e]) {
                double r0, r1, r2, om; int kd; Cam cam; Pose Trl;
                eval(i, T, r0, r1, r2, om, kd, cam, Trl);
                c = r0 * (om * r0) + r1 * (om * r1) + r2 * (om * r2);
                D.chi2[e] = c;
            }
            const float th = (kind == BAGPU_EDGE_STEREO) ? D.gate_stereo : D.gate_mono;
            const bool out = (float)c > th;
            D.outlier[e] = out ? 1 : 0;
            bad[0] += out ? 1.0 : 0.0;
        }
        block_allsum<1>(bad, sh);
        nBadGate = (int)(bad[0] + 0.5);
        if (it == 2) robust = false;
        if (n < 10) break;                                            // optimizer.edges().size() < 10
    }
    if (threadIdx.x == 0) {
        store_pose(D.pose_out + 7 * (size_t)f, T);
        D.n_inliers[f] = n - nBadGate;
        D.final_chi2[f] = lastChi;
    }
    (void)s_T; (void)s_flag;
}
