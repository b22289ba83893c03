// ba_ref -- CPU ORACLE for libbagpu.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call this file.  The product (libbagpu) never does.
//
// PARITY UNPINNED: the reference (hanseongbugi/ORB_SLAM3_Study_kr) ships no test,
// golden vector or fixture for its BA path and cannot be compiled
// here (Eigen3/OpenCV/Boost absent, Thirdparty/g2o/g2o/stuff/timeutil.h missing), so
// this file is a single-threaded FP64 *restatement* of the reference algorithm, file
// by file, and is pinned only by (a) the constants the reference source fixes
// (Huber deltas, chi2 gates, sigma table), (b) g2o's own numeric-Jacobian construction
// (base_binary_edge.hpp:130-205) and (c) an independent numpy/scipy solve (tests/).
//
// Reference files restated (paths relative to the ORB_SLAM3 tree):
//   Thirdparty/g2o/g2o/types/se3quat.h:53-64,104-128,217-285      SE3Quat
//   Thirdparty/g2o/g2o/types/se3_ops.hpp:27-38                    skew
//   Thirdparty/g2o/g2o/types/types_six_dof_expmap.{h,cpp}         VertexSE3Expmap, stereo edges
//   Thirdparty/g2o/g2o/types/types_sba.h:39-56                    VertexSBAPointXYZ
//   include/OptimizableTypes.h:31-144, src/OptimizableTypes.cpp   mono / body edges
//   src/CameraModels/Pinhole.cpp:35-41,71-81                      project / projectJac
//   src/CameraModels/KannalaBrandt8.cpp:46-65,145-175             project / projectJac
//   Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:66-91          Huber
//   Thirdparty/g2o/g2o/core/base_binary_edge.hpp:54-120           constructQuadraticForm
//   Thirdparty/g2o/g2o/core/base_unary_edge.hpp:42-72
//   Thirdparty/g2o/g2o/core/block_solver.hpp:143-295,353-486,501-604
//   Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:43-194
//   Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:61-114,166-267,354-435
//   Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124, linear_solver_dense.h:64-111
//   src/Optimizer.cc:61-390,815-1114,1116-1498,3506-3953          schedules and gates
// Eigen (un-vendored, version unpinned by the reference) is restated by its textbook
// formulas: quaternion-vector product, Hamilton product, Shepperd matrix->quaternion,
// adjugate 3x3 inverse, LDLT.
#include "../include/bagpu.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

// ---------------------------------------------------------------- small algebra
struct Q { double x, y, z, w; };
struct SE3 { Q q; double t[3]; };

inline void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// Eigen QuaternionBase::_transformVector: v + w*uv + q.vec x uv, uv = 2 (q.vec x v)
inline void qrot(const Q &q, const double v[3], double o[3]) {
    const double qv[3] = {q.x, q.y, q.z};
    double uv[3];
    cross3(qv, v, uv);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    double c[3];
    cross3(qv, uv, c);
    o[0] = v[0] + q.w * uv[0] + c[0];
    o[1] = v[1] + q.w * uv[1] + c[1];
    o[2] = v[2] + q.w * uv[2] + c[2];
}

// se3quat.h:280-285 normalizeRotation
inline void qnormalize(Q &q) {
    if (q.w < 0) { q.x *= -1; q.y *= -1; q.z *= -1; q.w *= -1; }
    const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

inline Q qmul(const Q &a, const Q &b) {
    Q r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}

// se3quat.h:104-110  operator*
inline SE3 se3_mul(const SE3 &a, const SE3 &b) {
    SE3 r = a;
    double rt[3];
    qrot(a.q, b.t, rt);
    r.t[0] += rt[0]; r.t[1] += rt[1]; r.t[2] += rt[2];
    r.q = qmul(a.q, b.q);
    qnormalize(r.q);
    return r;
}

// se3quat.h:217-220 map
inline void se3_map(const SE3 &T, const double X[3], double o[3]) {
    qrot(T.q, X, o);
    o[0] += T.t[0]; o[1] += T.t[1]; o[2] += T.t[2];
}

// Eigen toRotationMatrix (row-major R[r][c])
inline void qtoR(const Q &q, double R[3][3]) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz;       R[0][2] = txz + twy;
    R[1][0] = txy + twz;       R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
    R[2][0] = txz - twy;       R[2][1] = tyz + twx;       R[2][2] = 1 - (txx + tyy);
}

// Eigen Quaternion(Matrix3) -- Shepperd
inline Q qfromR(const double R[3][3]) {
    Q q;
    double t = R[0][0] + R[1][1] + R[2][2];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[2][1] - R[1][2]) * t;
        q.y = (R[0][2] - R[2][0]) * t;
        q.z = (R[1][0] - R[0][1]) * t;
    } else {
        int i = 0;
        if (R[1][1] > R[0][0]) i = 1;
        if (R[2][2] > R[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (R[k][j] - R[j][k]) * t;
        v[j] = (R[j][i] + R[i][j]) * t;
        v[k] = (R[k][i] + R[i][k]) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}

// se3quat.h:223-257 exp; update = [omega(3), upsilon(3)]
inline SE3 se3_exp(const double u[6]) {
    const double om[3] = {u[0], u[1], u[2]};
    const double up[3] = {u[3], u[4], u[5]};
    const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    double O[3][3] = {{0, -om[2], om[1]}, {om[2], 0, -om[0]}, {-om[1], om[0], 0}};
    double O2[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            O2[r][c] = O[r][0] * O[0][c] + O[r][1] * O[1][c] + O[r][2] * O[2][c];
    double R[3][3], V[3][3];
    if (theta < 0.00001) {
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                R[r][c] = ((r == c) ? 1.0 : 0.0) + O[r][c] + O2[r][c];
                V[r][c] = R[r][c];
            }
    } else {
        const double a = std::sin(theta) / theta;
        const double b = (1 - std::cos(theta)) / (theta * theta);
        const double c3 = (theta - std::sin(theta)) / (std::pow(theta, 3));
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                const double I = (r == c) ? 1.0 : 0.0;
                R[r][c] = I + a * O[r][c] + b * O2[r][c];
                V[r][c] = I + b * O[r][c] + c3 * O2[r][c];
            }
    }
    SE3 T;
    T.q = qfromR(R);
    for (int r = 0; r < 3; r++) T.t[r] = V[r][0] * up[0] + V[r][1] * up[1] + V[r][2] * up[2];
    qnormalize(T.q);   // SE3Quat(const Quaterniond&, const Vector3d&) ctor
    return T;
}

inline SE3 se3_from_qt(const double *qt) {
    SE3 T;
    T.t[0] = qt[0]; T.t[1] = qt[1]; T.t[2] = qt[2];
    T.q.x = qt[3]; T.q.y = qt[4]; T.q.z = qt[5]; T.q.w = qt[6];
    qnormalize(T.q);
    return T;
}
inline void se3_to_qt(const SE3 &T, double *qt) {
    qt[0] = T.t[0]; qt[1] = T.t[1]; qt[2] = T.t[2];
    qt[3] = T.q.x; qt[4] = T.q.y; qt[5] = T.q.z; qt[6] = T.q.w;
}

// ---------------------------------------------------------------- camera models
// Pinhole.cpp:35-41 / KannalaBrandt8.cpp:46-65 (float parameters, atan2f/sqrtf fossils)
inline void cam_project(const bagpu_camera &c, const double X[3], double uv[2]) {
    if (c.type == BAGPU_CAM_PINHOLE) {
        uv[0] = c.p[0] * X[0] / X[2] + c.p[2];
        uv[1] = c.p[1] * X[1] / X[2] + c.p[3];
    } else {
        const double x2_plus_y2 = X[0] * X[0] + X[1] * X[1];
        const double theta = atan2f(sqrtf(x2_plus_y2), X[2]);
        const double psi = atan2f(X[1], X[0]);
        const double theta2 = theta * theta;
        const double theta3 = theta * theta2;
        const double theta5 = theta3 * theta2;
        const double theta7 = theta5 * theta2;
        const double theta9 = theta7 * theta2;
        const double r = theta + c.p[4] * theta3 + c.p[5] * theta5 + c.p[6] * theta7 + c.p[7] * theta9;
        uv[0] = c.p[0] * r * cos(psi) + c.p[2];
        uv[1] = c.p[1] * r * sin(psi) + c.p[3];
    }
}

// Pinhole.cpp:71-81 / KannalaBrandt8.cpp:145-175; J[2][3]
inline void cam_project_jac(const bagpu_camera &c, const double X[3], double J[2][3]) {
    if (c.type == BAGPU_CAM_PINHOLE) {
        J[0][0] = c.p[0] / X[2];
        J[0][1] = 0.f;
        J[0][2] = -c.p[0] * X[0] / (X[2] * X[2]);
        J[1][0] = 0.f;
        J[1][1] = c.p[1] / X[2];
        J[1][2] = -c.p[1] * X[1] / (X[2] * X[2]);
    } else {
        double x2 = X[0] * X[0], y2 = X[1] * X[1], z2 = X[2] * X[2];
        double r2 = x2 + y2;
        double r = sqrt(r2);
        double r3 = r2 * r;
        double theta = atan2(r, X[2]);
        double theta2 = theta * theta, theta3 = theta2 * theta;
        double theta4 = theta2 * theta2, theta5 = theta4 * theta;
        double theta6 = theta2 * theta4, theta7 = theta6 * theta;
        double theta8 = theta4 * theta4, theta9 = theta8 * theta;
        double f = theta + theta3 * c.p[4] + theta5 * c.p[5] + theta7 * c.p[6] + theta9 * c.p[7];
        // "3 * mvParameters[4]" is an int*float product evaluated in float in the reference
        double fd = 1 + 3 * c.p[4] * theta2 + 5 * c.p[5] * theta4 + 7 * c.p[6] * theta6 + 9 * c.p[7] * theta8;
        J[0][0] = c.p[0] * (fd * X[2] * x2 / (r2 * (r2 + z2)) + f * y2 / r3);
        J[1][0] = c.p[1] * (fd * X[2] * X[1] * X[0] / (r2 * (r2 + z2)) - f * X[1] * X[0] / r3);
        J[0][1] = c.p[0] * (fd * X[2] * X[1] * X[0] / (r2 * (r2 + z2)) - f * X[1] * X[0] / r3);
        J[1][1] = c.p[1] * (fd * X[2] * y2 / (r2 * (r2 + z2)) + f * x2 / r3);
        J[0][2] = -c.p[0] * fd * X[0] / (r2 + z2);
        J[1][2] = -c.p[1] * fd * X[1] / (r2 + z2);
    }
}

// ---------------------------------------------------------------- edges
struct EdgeIn {
    int kind;                 // BAGPU_EDGE_*
    const bagpu_camera *cam;
    const SE3 *Trl;           // body only
    double obs[3];
};

// computeError(): OptimizableTypes.h:40-44,60-64,100-105,128-133; types_six_dof_expmap.h:155-160,.cpp:190-197,339-346
// unary==true selects the OnlyPose stereo variant, whose bf stays double (.cpp:344).
// Returns the error dimension (2 or 3).
inline int edge_error(const EdgeIn &e, const SE3 &Tcw, const double Xw[3], bool unary, double r[3]) {
    if (e.kind == BAGPU_EDGE_MONO) {
        double Xc[3], uv[2];
        se3_map(Tcw, Xw, Xc);
        cam_project(*e.cam, Xc, uv);
        r[0] = e.obs[0] - uv[0]; r[1] = e.obs[1] - uv[1]; r[2] = 0;
        return 2;
    } else if (e.kind == BAGPU_EDGE_BODY) {
        const SE3 Trw = se3_mul(*e.Trl, Tcw);     // (mTrl * v1->estimate()).map(X)
        double Xr[3], uv[2];
        se3_map(Trw, Xw, Xr);
        cam_project(*e.cam, Xr, uv);
        r[0] = e.obs[0] - uv[0]; r[1] = e.obs[1] - uv[1]; r[2] = 0;
        return 2;
    } else {
        double Xc[3];
        se3_map(Tcw, Xw, Xc);
        const double fx = e.cam->p[0], fy = e.cam->p[1], cx = e.cam->p[2], cy = e.cam->p[3];
        const float invz = 1.0f / Xc[2];
        double res0 = Xc[0] * invz * fx + cx;
        double res1 = Xc[1] * invz * fy + cy;
        double res2;
        if (unary) {
            const double bf = e.cam->bf;           // member double bf (set from float mbf)
            res2 = res0 - bf * invz;
        } else {
            const float bf = e.cam->bf;            // cam_project(..., const float &bf): float*float product
            res2 = res0 - bf * invz;
        }
        r[0] = e.obs[0] - res0; r[1] = e.obs[1] - res1; r[2] = e.obs[2] - res2;
        return 3;
    }
}

// isDepthPositive(): OptimizableTypes.h:46-49,66-69,107-111,135-139; types_six_dof_expmap.h:162-166
inline bool edge_depth_positive(const EdgeIn &e, const SE3 &Tcw, const double Xw[3]) {
    double Xc[3];
    if (e.kind == BAGPU_EDGE_BODY) {
        const SE3 Trw = se3_mul(*e.Trl, Tcw);
        se3_map(Trw, Xw, Xc);
    } else {
        se3_map(Tcw, Xw, Xc);
    }
    return Xc[2] > 0.0;
}

inline void se3deriv(const double X[3], double D[3][6]) {
    const double x = X[0], y = X[1], z = X[2];
    const double d[3][6] = {{0, z, -y, 1, 0, 0}, {-z, 0, x, 0, 1, 0}, {y, -x, 0, 0, 0, 1}};
    memcpy(D, d, sizeof(d));
}

// linearizeOplus(): A = d r / d point (dim x 3), B = d r / d pose (dim x 6)
// OptimizableTypes.cpp:49-63,91-107,139-160,192-213; types_six_dof_expmap.cpp:228-274,375-404
inline void edge_linearize(const EdgeIn &e, const SE3 &Tcw, const double Xw[3], bool unary,
                           double A[3][3], double B[3][6]) {
    memset(A, 0, sizeof(double) * 9);
    memset(B, 0, sizeof(double) * 18);
    if (e.kind == BAGPU_EDGE_MONO) {
        double Xc[3], J[2][3], R[3][3], D[3][6];
        se3_map(Tcw, Xw, Xc);
        cam_project_jac(*e.cam, Xc, J);
        for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) J[r][c] = -J[r][c];
        qtoR(Tcw.q, R);
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++) A[r][c] = J[r][0] * R[0][c] + J[r][1] * R[1][c] + J[r][2] * R[2][c];
        se3deriv(Xc, D);
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 6; c++) B[r][c] = J[r][0] * D[0][c] + J[r][1] * D[1][c] + J[r][2] * D[2][c];
    } else if (e.kind == BAGPU_EDGE_BODY) {
        const SE3 Trw = se3_mul(*e.Trl, Tcw);
        double Xl[3], Xr[3], J[2][3], R[3][3], Rrl[3][3], D[3][6];
        se3_map(Tcw, Xw, Xl);
        se3_map(*e.Trl, Xl, Xr);                   // mTrl.map(T_lw.map(X_w))
        cam_project_jac(*e.cam, Xr, J);
        for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) J[r][c] = -J[r][c];
        qtoR(Trw.q, R);
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++) A[r][c] = J[r][0] * R[0][c] + J[r][1] * R[1][c] + J[r][2] * R[2][c];
        qtoR(e.Trl->q, Rrl);
        double JR[2][3];
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++) JR[r][c] = J[r][0] * Rrl[0][c] + J[r][1] * Rrl[1][c] + J[r][2] * Rrl[2][c];
        se3deriv(Xl, D);
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 6; c++) B[r][c] = JR[r][0] * D[0][c] + JR[r][1] * D[1][c] + JR[r][2] * D[2][c];
    } else {
        double Xc[3];
        se3_map(Tcw, Xw, Xc);
        const double fx = e.cam->p[0], fy = e.cam->p[1], bf = e.cam->bf;
        const double x = Xc[0], y = Xc[1], z = Xc[2];
        if (!unary) {
            double R[3][3];
            qtoR(Tcw.q, R);
            const double z_2 = z * z;
            A[0][0] = -fx * R[0][0] / z + fx * x * R[2][0] / z_2;
            A[0][1] = -fx * R[0][1] / z + fx * x * R[2][1] / z_2;
            A[0][2] = -fx * R[0][2] / z + fx * x * R[2][2] / z_2;
            A[1][0] = -fy * R[1][0] / z + fy * y * R[2][0] / z_2;
            A[1][1] = -fy * R[1][1] / z + fy * y * R[2][1] / z_2;
            A[1][2] = -fy * R[1][2] / z + fy * y * R[2][2] / z_2;
            A[2][0] = A[0][0] - bf * R[2][0] / z_2;
            A[2][1] = A[0][1] - bf * R[2][1] / z_2;
            A[2][2] = A[0][2] - bf * R[2][2] / z_2;
            B[0][0] = x * y / z_2 * fx;
            B[0][1] = -(1 + (x * x / z_2)) * fx;
            B[0][2] = y / z * fx;
            B[0][3] = -1. / z * fx;
            B[0][4] = 0;
            B[0][5] = x / z_2 * fx;
            B[1][0] = (1 + y * y / z_2) * fy;
            B[1][1] = -x * y / z_2 * fy;
            B[1][2] = -x / z * fy;
            B[1][3] = 0;
            B[1][4] = -1. / z * fy;
            B[1][5] = y / z_2 * fy;
            B[2][0] = B[0][0] - bf * y / z_2;
            B[2][1] = B[0][1] + bf * x / z_2;
            B[2][2] = B[0][2];
            B[2][3] = B[0][3];
            B[2][4] = 0;
            B[2][5] = B[0][5] - bf / z_2;
        } else {
            const double invz = 1.0 / z;
            const double invz_2 = invz * invz;
            B[0][0] = x * y * invz_2 * fx;
            B[0][1] = -(1 + (x * x * invz_2)) * fx;
            B[0][2] = y * invz * fx;
            B[0][3] = -invz * fx;
            B[0][4] = 0;
            B[0][5] = x * invz_2 * fx;
            B[1][0] = (1 + y * y * invz_2) * fy;
            B[1][1] = -x * y * invz_2 * fy;
            B[1][2] = -x * invz * fy;
            B[1][3] = 0;
            B[1][4] = -invz * fy;
            B[1][5] = y * invz_2 * fy;
            B[2][0] = B[0][0] - bf * y * invz_2;
            B[2][1] = B[0][1] + bf * x * invz_2;
            B[2][2] = B[0][2];
            B[2][3] = B[0][3];
            B[2][4] = 0;
            B[2][5] = B[0][5] - bf * invz_2;
        }
    }
}

// robust_kernel_impl.cpp:78-91 (rho[2] unused: base_edge.h:98-104 comments the 2nd-order term out)
inline void huber(double e, double delta, double &rho0, double &rho1) {
    const double dsqr = delta * delta;
    if (e <= dsqr) { rho0 = e; rho1 = 1.; }
    else {
        const double sqrte = std::sqrt(e);
        rho0 = 2 * sqrte * delta - dsqr;
        rho1 = delta / sqrte;
    }
}

// Eigen fixed 3x3 inverse (cofactors * 1/det)
inline bool inv3(const double M[3][3], double I[3][3]) {
    const double c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1];
    const double c10 = M[1][2] * M[2][0] - M[1][0] * M[2][2];   // cofactor(1,0) -> goes to I[0][1]
    const double c20 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
    const double det = M[0][0] * c00 + M[0][1] * c10 + M[0][2] * c20;
    const double id = 1.0 / det;
    I[0][0] = c00 * id;
    I[1][0] = c10 * id;
    I[2][0] = c20 * id;
    I[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * id;
    I[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * id;
    I[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * id;
    I[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * id;
    I[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * id;
    I[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * id;
    return true;
}

// ---------------------------------------------------------------- linear solvers
// Reduced camera system, linear_solver_eigen.h:94-124: SimplicialLDLT<Upper>, no numerical
// pivoting, fails only on an exactly-zero pivot.  Restated as an envelope (skyline) LDL^T
// in natural order; the fill-reducing ordering (AMD) changes rounding only.
struct Skyline {
    int n = 0;
    std::vector<int> first;        // first[j]: first stored row of column j (upper triangle)
    std::vector<int64_t> colptr;   // column j stored at val[colptr[j] + (i-first[j])], i in [first[j], j]
    std::vector<double> val;
    void set_pattern(int n_, const std::vector<int> &first_) {
        n = n_; first = first_;
        colptr.assign(n + 1, 0);
        for (int j = 0; j < n; j++) colptr[j + 1] = colptr[j] + (j - first[j] + 1);
        val.assign(colptr[n], 0.0);
    }
    inline double &at(int i, int j) { return val[colptr[j] + (i - first[j])]; }   // i<=j, i>=first[j]
    // in-place A = U^T D U  (U unit upper); returns false on zero pivot
    bool factor() {
        std::vector<double> w;
        for (int j = 0; j < n; j++) {
            const int fj = first[j];
            double *cj = &val[colptr[j]] - fj;       // cj[i] = A(i,j)
            // solve for column j of U*D: v_i = A(i,j) - sum_{k<i} U(k,i) v_k  (v = D .* U(:,j))
            for (int i = fj; i < j; i++) {
                const int fi = first[i];
                const double *ci = &val[colptr[i]] - fi;
                const int k0 = std::max(fi, fj);
                double s = cj[i];
                for (int k = k0; k < i; k++) s -= ci[k] * cj[k];
                cj[i] = s;                            // v_i  (still multiplied by D)
            }
            double d = cj[j];
            for (int i = fj; i < j; i++) {
                const double di = val[colptr[i] + (i - first[i])];
                const double u = cj[i] / di;
                d -= cj[i] * u;
                cj[i] = u;
            }
            cj[j] = d;
            if (d == 0.0) return false;
        }
        return true;
    }
    void solve(const double *b, double *x) const {
        for (int i = 0; i < n; i++) x[i] = b[i];
        for (int j = 0; j < n; j++) {                 // U^T y = b
            const double *cj = &val[colptr[j]] - first[j];
            double s = x[j];
            for (int i = first[j]; i < j; i++) s -= cj[i] * x[i];
            x[j] = s;
        }
        for (int j = 0; j < n; j++) x[j] /= val[colptr[j] + (j - first[j])];
        for (int j = n - 1; j >= 0; j--) {            // U x = y
            const double *cj = &val[colptr[j]] - first[j];
            const double xj = x[j];
            for (int i = first[j]; i < j; i++) x[i] -= cj[i] * xj;
        }
    }
};

// linear_solver_dense.h:64-111: Eigen::LDLT (diagonal pivoting), isPositive() else fail.
inline bool dense_ldlt_solve(int n, const double *H /*row-major n*n symmetric*/, const double *b, double *x) {
    std::vector<double> A(H, H + (size_t)n * n);
    std::vector<int> perm(n);
    for (int i = 0; i < n; i++) perm[i] = i;
    bool positive = true;
    for (int k = 0; k < n; k++) {
        int p = k; double best = std::fabs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; i++) {
            const double v = std::fabs(A[(size_t)i * n + i]);
            if (v > best) { best = v; p = i; }
        }
        if (p != k) {
            for (int c = 0; c < n; c++) std::swap(A[(size_t)k * n + c], A[(size_t)p * n + c]);
            for (int r = 0; r < n; r++) std::swap(A[(size_t)r * n + k], A[(size_t)r * n + p]);
            std::swap(perm[k], perm[p]);
        }
        const double akk = A[(size_t)k * n + k];
        if (akk < 0) positive = false;
        if (akk == 0.0) { positive = positive && true; continue; }
        for (int i = k + 1; i < n; i++) A[(size_t)i * n + k] /= akk;           // L(i,k)
        for (int i = k + 1; i < n; i++) {
            const double lik = A[(size_t)i * n + k];
            for (int j = k + 1; j <= i; j++) {
                A[(size_t)i * n + j] -= lik * akk * A[(size_t)j * n + k];
                A[(size_t)j * n + i] = A[(size_t)i * n + j];
            }
        }
    }
    if (!positive) return false;
    std::vector<double> y(n);
    for (int i = 0; i < n; i++) y[i] = b[perm[i]];
    for (int i = 0; i < n; i++) { double s = y[i]; for (int k = 0; k < i; k++) s -= A[(size_t)i * n + k] * y[k]; y[i] = s; }
    for (int i = 0; i < n; i++) { const double d = A[(size_t)i * n + i]; y[i] = (d != 0.0) ? y[i] / d : 0.0; }
    for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= A[(size_t)k * n + i] * y[k]; y[i] = s; }
    for (int i = 0; i < n; i++) x[perm[i]] = y[i];
    return true;
}

// ---------------------------------------------------------------- BA problem state
struct Engine {
    const bagpu_problem *p;
    const bagpu_schedule *s;
    int Np, Nt;                          // points, total poses
    int64_t Ne;
    std::vector<SE3> pose, pose_bak, pose_init;
    std::vector<double> pt, pt_bak;      // 3*Np
    std::vector<SE3> rigs;
    std::vector<uint8_t> level, robust;
    std::vector<double> err;             // 3*Ne: e->_error
    // structure of the current optimize() (rebuilt per round; block_solver.hpp:143-295)
    std::vector<int> pose_h;             // pose -> hessian index among active free poses, else -1
    std::vector<int> pt_h;               // point -> landmark index among active points, else -1
    std::vector<int64_t> act;            // active edge ids in insertion order
    int Ncf = 0, Nl = 0;                 // active free poses / landmarks
    std::vector<int64_t> lm_ptr;         // CSR: per landmark, the active edges with a free pose (Hpl column)
    std::vector<int64_t> lm_edges;
    std::vector<double> Hpp_diag;        // Ncf * 36 (row-major 6x6, full)
    std::vector<double> Hll;             // Nl * 9
    std::vector<double> W;               // per (landmark,pose) block: 18 doubles (6x3 row-major)
    std::vector<int64_t> edge_w;         // edge -> W block index or -1
    std::vector<int64_t> w_ptr;          // per landmark: range of W blocks, sorted by pose h index
    std::vector<int> w_pose;             // pose h index per W block
    std::vector<double> b, x;            // 6*Ncf + 3*Nl
    Skyline S;
    std::vector<int> first_blk;          // per pose block column: first block row in Hschur pattern
    std::vector<double> coeff, bschur, Dinv;
    // timing / counters for the CPU baseline
    int64_t n_lin = 0, n_eval = 0, n_iter = 0, n_trial = 0;

    EdgeIn edge_in(int64_t e) const {
        EdgeIn in;
        in.kind = p->obs_kind[e];
        in.cam = &p->cameras[p->obs_cam[e]];
        in.Trl = (in.kind == BAGPU_EDGE_BODY) ? &rigs[p->obs_rig[e]] : nullptr;
        in.obs[0] = p->obs_u[e]; in.obs[1] = p->obs_v[e];
        in.obs[2] = (in.kind == BAGPU_EDGE_STEREO) ? p->obs_ur[e] : 0.0;
        return in;
    }
    inline double edge_chi2(int64_t e) const {   // base_edge.h:59-62 with information = inv_sigma2 * I
        const double *r = &err[3 * e];
        const double w = p->obs_inv_sigma2[e];
        return r[0] * (w * r[0]) + r[1] * (w * r[1]) + r[2] * (w * r[2]);
    }
    inline double delta_of(int64_t e) const { return p->obs_kind[e] == BAGPU_EDGE_STEREO ? s->delta_stereo : s->delta_mono; }

    void compute_error(int64_t e) {
        const EdgeIn in = edge_in(e);
        edge_error(in, pose[p->obs_pose[e]], &pt[3 * (size_t)p->obs_point[e]], false, &err[3 * e]);
    }
    // sparse_optimizer.cpp:61-76
    void compute_active_errors() { for (int64_t e : act) compute_error(e); n_eval += (int64_t)act.size(); }
    // sparse_optimizer.cpp:99-114
    double active_robust_chi2() const {
        double chi = 0.0;
        for (int64_t e : act) {
            const double c = edge_chi2(e);
            if (robust[e]) { double r0, r1; huber(c, delta_of(e), r0, r1); chi += r0; }
            else chi += c;
        }
        return chi;
    }

    // sparse_optimizer.cpp:199-267 + block_solver.hpp:143-295
    bool initialize(int lvl) {
        act.clear();
        std::vector<uint8_t> pose_act(Nt, 0), pt_act(Np, 0);
        for (int64_t e = 0; e < Ne; e++) {
            if (level[e] != lvl) continue;
            // allVerticesFixed(): the point is never fixed -> every level-matching edge is active
            act.push_back(e);
            pose_act[p->obs_pose[e]] = 1;
            pt_act[p->obs_point[e]] = 1;
        }
        pose_h.assign(Nt, -1); pt_h.assign(Np, -1);
        Ncf = 0; Nl = 0;
        for (int i = 0; i < Nt; i++) if (pose_act[i] && !p->pose_fixed[i]) pose_h[i] = Ncf++;
        for (int j = 0; j < Np; j++) if (pt_act[j]) pt_h[j] = Nl++;
        if (Ncf + Nl == 0) return false;
        // Hpl blocks: one per distinct (free pose, landmark) pair
        std::vector<std::vector<std::pair<int, int64_t>>> per_lm(Nl);   // (pose h, edge)
        for (int64_t e : act) {
            const int hp = pose_h[p->obs_pose[e]];
            if (hp < 0) continue;
            per_lm[pt_h[p->obs_point[e]]].push_back({hp, e});
        }
        edge_w.assign(Ne, -1);
        w_ptr.assign(Nl + 1, 0); w_pose.clear();
        first_blk.assign(Ncf, 0);
        for (int i = 0; i < Ncf; i++) first_blk[i] = i;
        for (int l = 0; l < Nl; l++) {
            auto &v = per_lm[l];
            std::stable_sort(v.begin(), v.end(), [](const std::pair<int, int64_t> &a, const std::pair<int, int64_t> &b) { return a.first < b.first; });
            int last = -1;
            for (auto &pe : v) {
                if (pe.first != last) { w_pose.push_back(pe.first); last = pe.first; }
                edge_w[pe.second] = (int64_t)w_pose.size() - 1;
            }
            w_ptr[l + 1] = (int64_t)w_pose.size();
            // Schur pattern: all pairs (i1<=i2) of this landmark's free poses (block_solver.hpp:262-288)
            const int64_t b0 = w_ptr[l], b1 = w_ptr[l + 1];
            if (b1 > b0) {
                const int lo = w_pose[b0];
                for (int64_t k = b0; k < b1; k++) first_blk[w_pose[k]] = std::min(first_blk[w_pose[k]], lo);
            }
        }
        W.assign(w_pose.size() * 18, 0.0);
        Hpp_diag.assign((size_t)Ncf * 36, 0.0);
        Hll.assign((size_t)Nl * 9, 0.0);
        b.assign((size_t)6 * Ncf + (size_t)3 * Nl, 0.0);
        x.assign(b.size(), 0.0);
        coeff.assign((size_t)6 * Ncf, 0.0); bschur.assign((size_t)6 * Ncf, 0.0);
        Dinv.assign((size_t)Nl * 9, 0.0);
        std::vector<int> first(6 * Ncf);
        for (int i = 0; i < Ncf; i++) for (int r = 0; r < 6; r++) first[6 * i + r] = 6 * first_blk[i];
        S.set_pattern(6 * Ncf, first);
        return true;
    }

    // block_solver.hpp:501-560 buildSystem: linearizeOplus + constructQuadraticForm per active edge
    void build_system() {
        std::fill(Hpp_diag.begin(), Hpp_diag.end(), 0.0);
        std::fill(Hll.begin(), Hll.end(), 0.0);
        std::fill(W.begin(), W.end(), 0.0);
        std::fill(b.begin(), b.end(), 0.0);
        for (int64_t e : act) {
            const EdgeIn in = edge_in(e);
            const int ip = p->obs_pose[e], jp = p->obs_point[e];
            double A[3][3], B[3][6];
            edge_linearize(in, pose[ip], &pt[3 * (size_t)jp], false, A, B);
            const int dim = (in.kind == BAGPU_EDGE_STEREO) ? 3 : 2;
            const double om = p->obs_inv_sigma2[e];
            const double *r = &err[3 * e];
            double rho1 = 1.0;
            if (robust[e]) { double r0; huber(edge_chi2(e), delta_of(e), r0, rho1); }
            double omega_r[3];
            for (int d = 0; d < dim; d++) omega_r[d] = -(om * r[d]);
            if (robust[e]) for (int d = 0; d < dim; d++) omega_r[d] *= rho1;
            const double wom = robust[e] ? rho1 * om : om;
            const int l = pt_h[jp];
            const int hp = pose_h[ip];
            double *bl = &b[(size_t)6 * Ncf + (size_t)3 * l];
            double *Hl = &Hll[(size_t)9 * l];
            for (int a = 0; a < 3; a++) {
                double s = 0; for (int d = 0; d < dim; d++) s += A[d][a] * omega_r[d];
                bl[a] += s;
                for (int c = 0; c < 3; c++) {
                    double h = 0; for (int d = 0; d < dim; d++) h += A[d][a] * wom * A[d][c];
                    Hl[3 * a + c] += h;
                }
            }
            if (hp >= 0) {
                double *bp = &b[(size_t)6 * hp];
                double *Hp = &Hpp_diag[(size_t)36 * hp];
                double *Wb = &W[(size_t)18 * edge_w[e]];
                for (int a = 0; a < 6; a++) {
                    double s = 0; for (int d = 0; d < dim; d++) s += B[d][a] * omega_r[d];
                    bp[a] += s;
                    for (int c = 0; c < 6; c++) {
                        double h = 0; for (int d = 0; d < dim; d++) h += B[d][a] * wom * B[d][c];
                        Hp[6 * a + c] += h;
                    }
                    for (int c = 0; c < 3; c++) {
                        double h = 0; for (int d = 0; d < dim; d++) h += B[d][a] * wom * A[d][c];
                        Wb[3 * a + c] += h;      // Hpl(pose, landmark), 6x3
                    }
                }
            }
        }
        n_lin += (int64_t)act.size();
    }

    // optimization_algorithm_levenberg.cpp:171-185
    double lambda_init() const {
        if (s->lambda_init > 0) return s->lambda_init;
        double m = 0.;
        for (int i = 0; i < Ncf; i++) for (int j = 0; j < 6; j++) m = std::max(std::fabs(Hpp_diag[(size_t)36 * i + 7 * j]), m);
        for (int l = 0; l < Nl; l++) for (int j = 0; j < 3; j++) m = std::max(std::fabs(Hll[(size_t)9 * l + 4 * j]), m);
        return 1e-5 * m;
    }

    // block_solver.hpp:353-486 solve() with the lambda of setLambda() folded in (restoreDiagonal is implicit)
    bool solve(double lambda) {
        if (Nl == 0) {   // no marginalised vertex: plain solve (optimization_algorithm_with_hessian.cpp:50-73)
            return false;
        }
        std::fill(S.val.begin(), S.val.end(), 0.0);
        for (int i = 0; i < Ncf; i++)
            for (int r = 0; r < 6; r++)
                for (int c = r; c < 6; c++)
                    S.at(6 * i + r, 6 * i + c) = Hpp_diag[(size_t)36 * i + 6 * r + c] + ((r == c) ? lambda : 0.0);
        std::fill(coeff.begin(), coeff.end(), 0.0);
        const double *bl0 = &b[(size_t)6 * Ncf];
        for (int l = 0; l < Nl; l++) {
            double D[3][3], Di[3][3];
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) D[r][c] = Hll[(size_t)9 * l + 3 * r + c] + ((r == c) ? lambda : 0.0);
            inv3(D, Di);
            memcpy(&Dinv[(size_t)9 * l], Di, sizeof(Di));
            double db[3];
            for (int r = 0; r < 3; r++) db[r] = Di[r][0] * bl0[3 * l] + Di[r][1] * bl0[3 * l + 1] + Di[r][2] * bl0[3 * l + 2];
            for (int64_t k1 = w_ptr[l]; k1 < w_ptr[l + 1]; k1++) {
                const int i1 = w_pose[k1];
                const double *Bi = &W[(size_t)18 * k1];
                double BD[6][3];
                for (int r = 0; r < 6; r++)
                    for (int c = 0; c < 3; c++) BD[r][c] = Bi[3 * r] * Di[0][c] + Bi[3 * r + 1] * Di[1][c] + Bi[3 * r + 2] * Di[2][c];
                for (int r = 0; r < 6; r++) coeff[(size_t)6 * i1 + r] += Bi[3 * r] * db[0] + Bi[3 * r + 1] * db[1] + Bi[3 * r + 2] * db[2];
                for (int64_t k2 = k1; k2 < w_ptr[l + 1]; k2++) {
                    const int i2 = w_pose[k2];
                    const double *Bj = &W[(size_t)18 * k2];
                    for (int r = 0; r < 6; r++) {
                        const int c0 = (i1 == i2) ? r : 0;
                        for (int c = c0; c < 6; c++)
                            S.at(6 * i1 + r, 6 * i2 + c) -= BD[r][0] * Bj[3 * c] + BD[r][1] * Bj[3 * c + 1] + BD[r][2] * Bj[3 * c + 2];
                    }
                }
            }
        }
        for (int i = 0; i < 6 * Ncf; i++) bschur[i] = b[i] - coeff[i];
        if (Ncf > 0) {
            if (!S.factor()) return false;
            S.solve(bschur.data(), x.data());
        }
        // back-substitution (block_solver.hpp:459-483): xl = Dinv (bl - W^T xp)
        double *xl0 = &x[(size_t)6 * Ncf];
        for (int l = 0; l < Nl; l++) {
            double cl[3] = {bl0[3 * l], bl0[3 * l + 1], bl0[3 * l + 2]};
            for (int64_t k = w_ptr[l]; k < w_ptr[l + 1]; k++) {
                const double *Bi = &W[(size_t)18 * k];
                const double *xp = &x[(size_t)6 * w_pose[k]];
                for (int c = 0; c < 3; c++) {
                    double sacc = 0; for (int r = 0; r < 6; r++) sacc += Bi[3 * r + c] * xp[r];
                    cl[c] -= sacc;
                }
            }
            const double *Di = &Dinv[(size_t)9 * l];
            for (int r = 0; r < 3; r++) xl0[3 * l + r] = Di[3 * r] * cl[0] + Di[3 * r + 1] * cl[1] + Di[3 * r + 2] * cl[2];
        }
        return true;
    }

    // sparse_optimizer.cpp:421-435 update(): oplus on every vertex of the index mapping
    void update() {
        for (int i = 0; i < Nt; i++) if (pose_h[i] >= 0) {
            const SE3 d = se3_exp(&x[(size_t)6 * pose_h[i]]);
            pose[i] = se3_mul(d, pose[i]);                 // types_six_dof_expmap.h:73-76
        }
        for (int j = 0; j < Np; j++) if (pt_h[j] >= 0) {
            const double *d = &x[(size_t)6 * Ncf + (size_t)3 * pt_h[j]];
            pt[3 * (size_t)j] += d[0]; pt[3 * (size_t)j + 1] += d[1]; pt[3 * (size_t)j + 2] += d[2];
        }
    }

    bool stop() const { return s->stop_flag && *s->stop_flag; }

    // one optimize(iterations) call: sparse_optimizer.cpp:354-419 + optimization_algorithm_levenberg.cpp:61-169
    int optimize(int round, int iterations, bagpu_result *res) {
        int status = BAGPU_OK;
        double lambda = -1, ni = 2; int nBad = 0;
        bool ok = true;
        for (int it = 0; it < iterations && !stop() && ok; it++) {
            compute_active_errors();
            double currentChi = active_robust_chi2();
            double tempChi = currentChi;
            const double iniChi = currentChi;
            build_system();
            if (it == 0) { lambda = lambda_init(); ni = 2; nBad = 0; }
            double rho = 0; int qmax = 0;
            do {
                pose_bak = pose; pt_bak = pt;                       // push()
                const bool ok2 = solve(lambda);
                update();
                compute_active_errors();
                tempChi = active_robust_chi2();
                if (!ok2) tempChi = DBL_MAX;
                rho = (currentChi - tempChi);
                double scale = 0.;
                for (size_t j = 0; j < x.size(); j++) scale += x[j] * (lambda * x[j] + b[j]);
                scale += 1e-3;
                rho /= scale;
                if (rho > 0 && std::isfinite(tempChi)) {
                    double alpha = 1. - std::pow((2 * rho - 1), 3);
                    alpha = std::min(alpha, 2. / 3.);
                    const double scaleFactor = std::max(1. / 3., alpha);
                    lambda *= scaleFactor;
                    ni = 2;
                    currentChi = tempChi;
                } else {
                    lambda *= ni;
                    ni *= 2;
                    pose = pose_bak; pt = pt_bak;                   // pop(); edge errors are NOT recomputed
                }
                qmax++; n_trial++;
            } while (rho < 0 && qmax < 10 && !stop());
            n_iter++;
            int st = BAGPU_OK;
            if (qmax == 10 || rho == 0) st = BAGPU_TERMINATE_TRIALS;
            else {
                if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
                if (nBad >= 3) st = BAGPU_TERMINATE_NBAD;
            }
            if (res && res->trace && res->n_trace < s->max_trace) {
                bagpu_trace &t = res->trace[res->n_trace++];
                t.round = round; t.iteration = it; t.chi2_before = iniChi; t.chi2_after = currentChi;
                t.lambda = lambda; t.trials = qmax; t.status = st;
                t.active_edges = 0; t.linearise_schur_us = t.linear_solve_us = t.update_us = t.iteration_us = 0.0;   // phase record: GPU library only
            }
            status = st;
            ok = (st == BAGPU_OK);
        }
        if (stop() && status == BAGPU_OK) status = BAGPU_STOPPED;
        return status;
    }
};

}  // namespace

// ================================================================== C entry points
extern "C" {

// Same contract as bagpu_solve_ba (include/bagpu.h), computed on one CPU thread.
int ba_ref_solve(const bagpu_problem *p, const bagpu_schedule *s, bagpu_result *r, int64_t *counters /*[4] lin,eval,iter,trial or NULL*/) {
    Engine E;
    E.p = p; E.s = s; E.Np = p->n_points; E.Nt = p->n_poses; E.Ne = p->n_obs;
    E.pose.resize(E.Nt);
    for (int i = 0; i < E.Nt; i++) E.pose[i] = se3_from_qt(&p->pose_qt[7 * (size_t)i]);
    E.pose_init = E.pose;
    E.pt.assign(p->points, p->points + 3 * (size_t)E.Np);
    E.rigs.resize(p->n_rigs);
    for (int i = 0; i < p->n_rigs; i++) E.rigs[i] = se3_from_qt(p->rigs[i].qt);
    E.level.assign(E.Ne, 0);
    E.robust.resize(E.Ne);
    for (int64_t e = 0; e < E.Ne; e++) E.robust[e] = (p->obs_flags[e] & BAGPU_FLAG_ROBUST) ? 1 : 0;
    E.err.assign(3 * (size_t)E.Ne, 0.0);
    if (r) { r->n_trace = 0; r->status = BAGPU_OK; }
    int status = BAGPU_OK;
    for (int k = 0; k < s->n_rounds; k++) {
        const bagpu_round &rd = s->rounds[k];
        if (rd.reset_pose) E.pose = E.pose_init;
        if (E.stop()) { status = BAGPU_STOPPED; break; }     // Optimizer.cc:3725-3737 (bDoMore)
        if (E.initialize(0) && !E.act.empty()) status = E.optimize(k, rd.iterations, r);
        if (rd.gate_after == BAGPU_GATE_LBA) {                // Optimizer.cc:3745-3776
            if (E.stop()) { status = BAGPU_STOPPED; break; }
            for (int64_t e = 0; e < E.Ne; e++) {
                const double th = (p->obs_kind[e] == BAGPU_EDGE_STEREO) ? rd.gate_stereo : rd.gate_mono;
                const EdgeIn in = E.edge_in(e);
                const bool dp = edge_depth_positive(in, E.pose[p->obs_pose[e]], &E.pt[3 * (size_t)p->obs_point[e]]);
                if (E.edge_chi2(e) > th || !dp) E.level[e] = 1;
            }
        } else if (rd.gate_after == BAGPU_GATE_POSE) {         // Optimizer.cc:1013-1100
            for (int64_t e = 0; e < E.Ne; e++) {
                if (E.level[e] == 1) E.compute_error(e);
                const float chi2 = (float)E.edge_chi2(e);
                const float th = (p->obs_kind[e] == BAGPU_EDGE_STEREO) ? (float)rd.gate_stereo : (float)rd.gate_mono;
                E.level[e] = (chi2 > th) ? 1 : 0;
            }
        }
        if (rd.drop_kernel_after) std::fill(E.robust.begin(), E.robust.end(), 0);
    }
    if (r) {
        r->status = status;
        if (r->pose_qt) for (int i = 0; i < E.Nt; i++) se3_to_qt(E.pose[i], &r->pose_qt[7 * (size_t)i]);
        if (r->points) memcpy(r->points, E.pt.data(), sizeof(double) * 3 * (size_t)E.Np);
        for (int64_t e = 0; e < E.Ne; e++) {
            if (r->edge_chi2) r->edge_chi2[e] = E.edge_chi2(e);
            if (r->edge_depth_pos) {
                const EdgeIn in = E.edge_in(e);
                r->edge_depth_pos[e] = edge_depth_positive(in, E.pose[p->obs_pose[e]], &E.pt[3 * (size_t)p->obs_point[e]]) ? 1 : 0;
            }
            if (r->edge_level) r->edge_level[e] = E.level[e];
        }
    }
    if (counters) { counters[0] = E.n_lin; counters[1] = E.n_eval; counters[2] = E.n_iter; counters[3] = E.n_trial; }
    return status;
}

// Optimizer::PoseOptimization for a batch of independent frames (Optimizer.cc:815-1114).
int ba_ref_pose_opt_batch(const bagpu_pose_batch *B, bagpu_pose_result *R) {
    std::vector<SE3> rigs(B->n_rigs);
    for (int i = 0; i < B->n_rigs; i++) rigs[i] = se3_from_qt(B->rigs[i].qt);
    for (int f = 0; f < B->n_frames; f++) {
        const int64_t e0 = B->frame_ptr[f], e1 = B->frame_ptr[f + 1];
        const int n = (int)(e1 - e0);
        const SE3 T0 = se3_from_qt(&B->pose_qt[7 * (size_t)f]);
        SE3 T = T0;
        std::vector<uint8_t> level(n, 0), robust(n, 1);
        std::vector<double> err(3 * (size_t)n, 0.0);
        double lastChi = 0.0;
        for (int i = 0; i < n; i++) if (R->outlier) R->outlier[e0 + i] = 0;
        if (n < 3) {                                           // Optimizer.cc:996-997
            if (R->pose_qt) se3_to_qt(T0, &R->pose_qt[7 * (size_t)f]);
            if (R->n_inliers) R->n_inliers[f] = 0;
            if (R->final_chi2) R->final_chi2[f] = 0.0;
            continue;
        }
        auto mk = [&](int i) {
            EdgeIn in; const int64_t e = e0 + i;
            in.kind = B->obs_kind[e]; in.cam = &B->cameras[B->obs_cam[e]];
            in.Trl = (in.kind == BAGPU_EDGE_BODY) ? &rigs[B->obs_rig[e]] : nullptr;
            in.obs[0] = B->obs_u[e]; in.obs[1] = B->obs_v[e];
            in.obs[2] = (in.kind == BAGPU_EDGE_STEREO) ? B->obs_ur[e] : 0.0;
            return in;
        };
        auto chi2_of = [&](int i) {
            const double *r = &err[3 * (size_t)i]; const double w = B->obs_inv_sigma2[e0 + i];
            return r[0] * (w * r[0]) + r[1] * (w * r[1]) + r[2] * (w * r[2]);
        };
        auto delta_of = [&](int i) { return B->obs_kind[e0 + i] == BAGPU_EDGE_STEREO ? B->delta_stereo : B->delta_mono; };
        int nBad = 0;
        for (int it = 0; it < 4; it++) {
            T = T0;                                            // vSE3->setEstimate(initial pose)
            std::vector<int> act;
            for (int i = 0; i < n; i++) if (level[i] == 0) act.push_back(i);
            if (!act.empty()) {
                // optimize(10): unary edges, 6x6 system, LinearSolverDense
                double lambda = -1, ni = 2; int nb = 0; bool ok = true;
                auto errors = [&]() { for (int i : act) { const EdgeIn in = mk(i); edge_error(in, T, &B->xw[3 * (size_t)(e0 + i)], true, &err[3 * (size_t)i]); } };
                auto rchi2 = [&]() {
                    double chi = 0;
                    for (int i : act) { const double c = chi2_of(i); if (robust[i]) { double r0, r1; huber(c, delta_of(i), r0, r1); chi += r0; } else chi += c; }
                    return chi;
                };
                for (int k = 0; k < 10 && ok; k++) {
                    errors();
                    double currentChi = rchi2(), tempChi = currentChi; const double iniChi = currentChi;
                    double H[36], b[6], x[6] = {0, 0, 0, 0, 0, 0};
                    memset(H, 0, sizeof(H)); memset(b, 0, sizeof(b));
                    for (int i : act) {                        // base_unary_edge.hpp:42-72
                        const EdgeIn in = mk(i);
                        double A[3][3], Bm[3][6];
                        edge_linearize(in, T, &B->xw[3 * (size_t)(e0 + i)], true, A, Bm);
                        const int dim = (in.kind == BAGPU_EDGE_STEREO) ? 3 : 2;
                        const double om = B->obs_inv_sigma2[e0 + i];
                        const double *r = &err[3 * (size_t)i];
                        double rho1 = 1.0;
                        if (robust[i]) { double r0; huber(chi2_of(i), delta_of(i), r0, rho1); }
                        const double wom = robust[i] ? rho1 * om : om;
                        for (int a = 0; a < 6; a++) {
                            double sacc = 0; for (int d = 0; d < dim; d++) sacc += Bm[d][a] * om * r[d];
                            b[a] -= robust[i] ? rho1 * sacc : sacc;
                            for (int c = 0; c < 6; c++) { double h = 0; for (int d = 0; d < dim; d++) h += Bm[d][a] * wom * Bm[d][c]; H[6 * a + c] += h; }
                        }
                    }
                    if (k == 0) { double m = 0; for (int j = 0; j < 6; j++) m = std::max(std::fabs(H[7 * j]), m); lambda = 1e-5 * m; ni = 2; nb = 0; }
                    double rho = 0; int qmax = 0;
                    do {
                        const SE3 Tbak = T;
                        double Hl[36]; memcpy(Hl, H, sizeof(H));
                        for (int j = 0; j < 6; j++) Hl[7 * j] += lambda;
                        const bool ok2 = dense_ldlt_solve(6, Hl, b, x);
                        T = se3_mul(se3_exp(x), T);
                        errors();
                        tempChi = rchi2();
                        if (!ok2) tempChi = DBL_MAX;
                        rho = currentChi - tempChi;
                        double scale = 0; for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
                        scale += 1e-3; rho /= scale;
                        if (rho > 0 && std::isfinite(tempChi)) {
                            double alpha = 1. - std::pow((2 * rho - 1), 3);
                            alpha = std::min(alpha, 2. / 3.);
                            lambda *= std::max(1. / 3., alpha); ni = 2; currentChi = tempChi;
                        } else { lambda *= ni; ni *= 2; T = Tbak; }
                        qmax++;
                    } while (rho < 0 && qmax < 10);
                    lastChi = currentChi;
                    if (qmax == 10 || rho == 0) ok = false;
                    else { if ((iniChi - currentChi) * 1e3 < iniChi) nb++; else nb = 0; if (nb >= 3) ok = false; }
                }
            }
            nBad = 0;
            for (int i = 0; i < n; i++) {
                if (level[i] == 1) { const EdgeIn in = mk(i); edge_error(in, T, &B->xw[3 * (size_t)(e0 + i)], true, &err[3 * (size_t)i]); }
                const float chi2 = (float)chi2_of(i);
                const float th = (B->obs_kind[e0 + i] == BAGPU_EDGE_STEREO) ? B->gate_stereo : B->gate_mono;
                if (chi2 > th) { level[i] = 1; nBad++; } else level[i] = 0;
                if (it == 2) robust[i] = 0;
            }
            if (n < 10) break;                                  // optimizer.edges().size()<10
        }
        if (R->pose_qt) se3_to_qt(T, &R->pose_qt[7 * (size_t)f]);
        if (R->outlier) for (int i = 0; i < n; i++) R->outlier[e0 + i] = level[i];
        if (R->n_inliers) R->n_inliers[f] = n - nBad;
        if (R->final_chi2) R->final_chi2[f] = lastChi;
    }
    return 0;
}

// ---- unit hooks for tests (edge arithmetic, Lie group, solvers) -----------------
// kind, camera, optional Trl (7), pose (7), X (3), obs (3), unary -> err(3), A(9 row-major), B(18 row-major), depth_pos
void ba_ref_edge(int kind, const bagpu_camera *cam, const double *trl_qt, const double *pose_qt, const double *X,
                 const double *obs, int unary, double *err, double *A, double *Bm, int *depth_pos) {
    SE3 Trl; if (trl_qt) Trl = se3_from_qt(trl_qt);
    EdgeIn in; in.kind = kind; in.cam = cam; in.Trl = trl_qt ? &Trl : nullptr;
    in.obs[0] = obs[0]; in.obs[1] = obs[1]; in.obs[2] = obs[2];
    const SE3 T = se3_from_qt(pose_qt);
    edge_error(in, T, X, unary != 0, err);
    double a[3][3], b[3][6];
    edge_linearize(in, T, X, unary != 0, a, b);
    memcpy(A, a, sizeof(a)); memcpy(Bm, b, sizeof(b));
    *depth_pos = edge_depth_positive(in, T, X) ? 1 : 0;
}
// pose <- exp(update) * pose
void ba_ref_oplus(const double *pose_qt, const double *update6, double *out_qt) {
    const SE3 T = se3_mul(se3_exp(update6), se3_from_qt(pose_qt));
    se3_to_qt(T, out_qt);
}
void ba_ref_huber(double e, double delta, double *rho0, double *rho1) { huber(e, delta, *rho0, *rho1); }
void ba_ref_atan2f(const float *y, const float *x, float *out, int64_t n) { for (int64_t i = 0; i < n; i++) out[i] = atan2f(y[i], x[i]); }
int ba_ref_dense_ldlt(int n, const double *H, const double *b, double *x) { return dense_ldlt_solve(n, H, b, x) ? 1 : 0; }
// dense symmetric matrix -> skyline LDL^T solve (tests the reduced-system solver against numpy)
int ba_ref_skyline_solve(int n, const double *H, const double *b, double *x) {
    std::vector<int> first(n);
    for (int j = 0; j < n; j++) { int f = j; for (int i = 0; i < j; i++) if (H[(size_t)i * n + j] != 0.0) { f = i; break; } first[j] = f; }
    Skyline S; S.set_pattern(n, first);
    for (int j = 0; j < n; j++) for (int i = first[j]; i <= j; i++) S.at(i, j) = H[(size_t)i * n + j];
    if (!S.factor()) return 0;
    S.solve(b, x);
    return 1;
}

}  // extern "C"
