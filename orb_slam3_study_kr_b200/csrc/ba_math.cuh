// Device arithmetic of the BA hot path: SE3, camera models, edge residuals and Jacobians, Huber.
// What each function computes is fixed by the reference (cited per function); how it is organised
// (one "linearised edge" record per lane, camera-frame Jacobian factored out for all edge kinds) is ours.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/bagpu.h"
#include "fdlibm_atan2f.h"

#define BA_DEV __device__ __forceinline__

struct Quat { double x, y, z, w; };
struct Pose { double tx, ty, tz; Quat q; };

BA_DEV Pose load_pose(const double *__restrict__ p) {
    Pose T;
    T.tx = __ldg(p + 0); T.ty = __ldg(p + 1); T.tz = __ldg(p + 2);
    T.q.x = __ldg(p + 3); T.q.y = __ldg(p + 4); T.q.z = __ldg(p + 5); T.q.w = __ldg(p + 6);
    return T;
}
BA_DEV void store_pose(double *p, const Pose &T) {
    p[0] = T.tx; p[1] = T.ty; p[2] = T.tz; p[3] = T.q.x; p[4] = T.q.y; p[5] = T.q.z; p[6] = T.q.w;
}

// Eigen quaternion * vector as used by SE3Quat::map (Thirdparty/g2o/g2o/types/se3quat.h:217-220):
// v + w*uv + q.vec x uv with uv = 2 (q.vec x v)
BA_DEV void quat_rotate(const Quat &q, double vx, double vy, double vz, double &ox, double &oy, double &oz) {
    double ux = q.y * vz - q.z * vy, uy = q.z * vx - q.x * vz, uz = q.x * vy - q.y * vx;
    ux += ux; uy += uy; uz += uz;
    ox = vx + q.w * ux + (q.y * uz - q.z * uy);
    oy = vy + q.w * uy + (q.z * ux - q.x * uz);
    oz = vz + q.w * uz + (q.x * uy - q.y * ux);
}
BA_DEV void pose_map(const Pose &T, double X, double Y, double Z, double &ox, double &oy, double &oz) {
    quat_rotate(T.q, X, Y, Z, ox, oy, oz);
    ox += T.tx; oy += T.ty; oz += T.tz;
}
// se3quat.h:280-285 normalizeRotation
BA_DEV void quat_normalize(Quat &q) {
    if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}
BA_DEV Quat quat_mul(const Quat &a, const Quat &b) {
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
// se3quat.h:104-110 operator*
BA_DEV Pose pose_mul(const Pose &a, const Pose &b) {
    Pose r;
    double rx, ry, rz;
    quat_rotate(a.q, b.tx, b.ty, b.tz, rx, ry, rz);
    r.tx = a.tx + rx; r.ty = a.ty + ry; r.tz = a.tz + rz;
    r.q = quat_mul(a.q, b.q);
    quat_normalize(r.q);
    return r;
}
// Eigen toRotationMatrix, row-major R[3*r+c]
BA_DEV void quat_to_R(const Quat &q, double *R) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
// Eigen Quaternion(Matrix3): Shepperd's method
BA_DEV Quat quat_from_R(const double *R) {
    Quat q;
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[7] - R[5]) * t; q.y = (R[2] - R[6]) * t; q.z = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (R[3 * k + j] - R[3 * j + k]) * t;
        v[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        v[k] = (R[3 * k + i] + R[3 * i + k]) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}
// SE3Quat::exp (se3quat.h:223-257) followed by the left product of VertexSE3Expmap::oplusImpl
// (types_six_dof_expmap.h:73-76): T <- exp(u) * T, u = [omega, upsilon]
BA_DEV Pose pose_oplus(const Pose &T, const double *u) {
    const double ox = u[0], oy = u[1], oz = u[2];
    const double theta = sqrt(ox * ox + oy * oy + oz * oz);
    const double O[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
    double O2[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) O2[3 * r + c] = O[3 * r] * O[c] + O[3 * r + 1] * O[3 + c] + O[3 * r + 2] * O[6 + c];
    double R[9], V[9];
    if (theta < 0.00001) {
#pragma unroll
        for (int i = 0; i < 9; i++) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        double s, c;
        sincos(theta, &s, &c);
        const double a = s / theta, b = (1 - c) / (theta * theta), c3 = (theta - s) / (theta * theta * theta);
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const double I = (i % 4 == 0) ? 1.0 : 0.0;
            R[i] = I + a * O[i] + b * O2[i];
            V[i] = I + b * O[i] + c3 * O2[i];
        }
    }
    Pose d;
    d.q = quat_from_R(R);
    quat_normalize(d.q);
    d.tx = V[0] * u[3] + V[1] * u[4] + V[2] * u[5];
    d.ty = V[3] * u[3] + V[4] * u[4] + V[5] * u[5];
    d.tz = V[6] * u[3] + V[7] * u[4] + V[8] * u[5];
    return pose_mul(d, T);
}

// ------------------------------------------------------------------ cameras
struct Cam { int type; float p[8]; float bf; };
BA_DEV Cam load_cam(const bagpu_camera *__restrict__ c) {
    Cam k;
    k.type = c->type;
#pragma unroll
    for (int i = 0; i < 8; i++) k.p[i] = c->p[i];
    k.bf = c->bf;
    return k;
}

// Pinhole::project / KannalaBrandt8::project on a double point with float parameters
// (src/CameraModels/Pinhole.cpp:35-41, KannalaBrandt8.cpp:46-65). KB8 keeps the reference's
// float atan2f/sqrtf on narrowed arguments.
BA_DEV void cam_project(const Cam &c, double x, double y, double z, double &u, double &v) {
    if (c.type == BAGPU_CAM_PINHOLE) {
        u = (double)c.p[0] * x / z + (double)c.p[2];
        v = (double)c.p[1] * y / z + (double)c.p[3];
    } else {
        const double x2_plus_y2 = x * x + y * y;
        const double theta = (double)baf_atan2f(__fsqrt_rn(__double2float_rn(x2_plus_y2)), __double2float_rn(z));
        const double psi = (double)baf_atan2f(__double2float_rn(y), __double2float_rn(x));
        const double theta2 = theta * theta;
        const double theta3 = theta * theta2;
        const double theta5 = theta3 * theta2;
        const double theta7 = theta5 * theta2;
        const double theta9 = theta7 * theta2;
        const double r = theta + (double)c.p[4] * theta3 + (double)c.p[5] * theta5 + (double)c.p[6] * theta7 + (double)c.p[7] * theta9;
        double sp, cp;
        sincos(psi, &sp, &cp);
        u = (double)c.p[0] * r * cp + (double)c.p[2];
        v = (double)c.p[1] * r * sp + (double)c.p[3];
    }
}

// NEGATED projection Jacobian  Jn = -projectJac(X)  (2x3, row-major J[3*r+c])
// (Pinhole.cpp:71-81, KannalaBrandt8.cpp:145-175; "3 * mvParameters[4]" is a float product there)
BA_DEV void cam_neg_jac(const Cam &c, double x, double y, double z, double *J) {
    if (c.type == BAGPU_CAM_PINHOLE) {
        const double fx = (double)c.p[0], fy = (double)c.p[1];
        J[0] = -(fx / z); J[1] = -0.0; J[2] = -((double)(-c.p[0]) * x / (z * z));
        J[3] = -0.0; J[4] = -(fy / z); J[5] = -((double)(-c.p[1]) * y / (z * z));
    } else {
        const double x2 = x * x, y2 = y * y, z2 = z * z;
        const double r2 = x2 + y2;
        const double r = sqrt(r2);
        const double r3 = r2 * r;
        const double theta = atan2(r, z);
        const double theta2 = theta * theta, theta3 = theta2 * theta;
        const double theta4 = theta2 * theta2, theta5 = theta4 * theta;
        const double theta6 = theta2 * theta4, theta7 = theta6 * theta;
        const double theta8 = theta4 * theta4, theta9 = theta8 * theta;
        const double f = theta + theta3 * (double)c.p[4] + theta5 * (double)c.p[5] + theta7 * (double)c.p[6] + theta9 * (double)c.p[7];
        const double fd = 1 + (double)__fmul_rn(3.0f, c.p[4]) * theta2 + (double)__fmul_rn(5.0f, c.p[5]) * theta4 +
                          (double)__fmul_rn(7.0f, c.p[6]) * theta6 + (double)__fmul_rn(9.0f, c.p[7]) * theta8;
        const double fx = (double)c.p[0], fy = (double)c.p[1];
        const double den = r2 * (r2 + z2);
        const double j00 = fx * (fd * z * x2 / den + f * y2 / r3);
        const double j10 = fy * (fd * z * y * x / den - f * y * x / r3);
        const double j01 = fx * (fd * z * y * x / den - f * y * x / r3);
        const double j11 = fy * (fd * z * y2 / den + f * x2 / r3);
        const double j02 = (double)(-c.p[0]) * fd * x / (r2 + z2);
        const double j12 = (double)(-c.p[1]) * fd * y / (r2 + z2);
        J[0] = -j00; J[1] = -j01; J[2] = -j02; J[3] = -j10; J[4] = -j11; J[5] = -j12;
    }
}

// ------------------------------------------------------------------ edges
// Residual of one edge (computeError of the six edge classes):
//   mono   ORB_SLAM3::EdgeSE3ProjectXYZ(OnlyPose)        include/OptimizableTypes.h:40-44,100-105
//   body   ORB_SLAM3::EdgeSE3ProjectXYZ(OnlyPose)ToBody  include/OptimizableTypes.h:60-64,128-133
//   stereo g2o::EdgeStereoSE3ProjectXYZ(OnlyPose)        types_six_dof_expmap.cpp:190-197,339-346
// `unary` selects the OnlyPose stereo variant whose bf stays double. Returns depth (z in the edge's camera).
BA_DEV double edge_residual(int kind, const Cam &c, const Pose &T, const Pose *Trl, double X, double Y, double Z,
                            double ou, double ov, double our, bool unary, double &r0, double &r1, double &r2) {
    double x, y, z;
    if (kind == BAGPU_EDGE_BODY) {
        const Pose Trw = pose_mul(*Trl, T);
        pose_map(Trw, X, Y, Z, x, y, z);
    } else {
        pose_map(T, X, Y, Z, x, y, z);
    }
    if (kind == BAGPU_EDGE_STEREO) {
        const float invz = __double2float_rn(1.0 / z);
        const double fx = (double)c.p[0], fy = (double)c.p[1], cx = (double)c.p[2], cy = (double)c.p[3];
        const double pu = x * (double)invz * fx + cx;
        const double pv = y * (double)invz * fy + cy;
        const double disp = unary ? ((double)c.bf * (double)invz) : (double)__fmul_rn(c.bf, invz);
        r0 = ou - pu; r1 = ov - pv; r2 = our - (pu - disp);
    } else {
        double pu, pv;
        cam_project(c, x, y, z, pu, pv);
        r0 = ou - pu; r1 = ov - pv; r2 = 0.0;
    }
    return z;
}

// Linearised edge: A = dr/dpoint (dim x 3), B = dr/dpose (dim x 6, [rotation | translation]).
// All kinds share   A = Jn * R_total,   B = (Jn * R_rl) * [ -[X_l]x | I ]
// with Jn the negated projection Jacobian at the edge's camera-frame point (OptimizableTypes.cpp:49-63,
// 91-107,139-160,192-213). For stereo this is algebraically the closed form of
// types_six_dof_expmap.cpp:228-274,375-404 (same entries, different association of the products).
struct EdgeLin {
    double A[9];      // rows 0..dim-1 used
    double B[18];     // B = M [ -[X_l]x | I ]: columns 3..5 are M = Jn R_rl
    double xl[3];     // the point in the (left) camera frame
};

BA_DEV void edge_linearize(int kind, const Cam &c, const Pose &T, const Pose *Trl, double X, double Y, double Z, EdgeLin &L) {
    double xl, yl, zl;
    pose_map(T, X, Y, Z, xl, yl, zl);
    double Jn[9];                       // dim x 3
    double R[9];
    double M[9];                        // Jn * R_rl
    if (kind == BAGPU_EDGE_BODY) {
        double xr, yr, zr;
        pose_map(*Trl, xl, yl, zl, xr, yr, zr);            // mTrl.map(T_lw.map(X_w))
        cam_neg_jac(c, xr, yr, zr, Jn);
        Jn[6] = Jn[7] = Jn[8] = 0.0;
        const Pose Trw = pose_mul(*Trl, T);
        quat_to_R(Trw.q, R);
        double Rrl[9];
        quat_to_R(Trl->q, Rrl);
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int cc = 0; cc < 3; cc++) M[3 * r + cc] = Jn[3 * r] * Rrl[cc] + Jn[3 * r + 1] * Rrl[3 + cc] + Jn[3 * r + 2] * Rrl[6 + cc];
        M[6] = M[7] = M[8] = 0.0;
    } else {
        if (kind == BAGPU_EDGE_STEREO) {
            const double fx = (double)c.p[0], fy = (double)c.p[1], bf = (double)c.bf;
            const double iz = 1.0 / zl, iz2 = iz * iz;
            Jn[0] = -fx * iz; Jn[1] = 0.0; Jn[2] = fx * xl * iz2;
            Jn[3] = 0.0; Jn[4] = -fy * iz; Jn[5] = fy * yl * iz2;
            Jn[6] = Jn[0]; Jn[7] = 0.0; Jn[8] = Jn[2] - bf * iz2;
        } else {
            cam_neg_jac(c, xl, yl, zl, Jn);
            Jn[6] = Jn[7] = Jn[8] = 0.0;
        }
        quat_to_R(T.q, R);
#pragma unroll
        for (int i = 0; i < 9; i++) M[i] = Jn[i];
    }
    L.xl[0] = xl; L.xl[1] = yl; L.xl[2] = zl;
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int cc = 0; cc < 3; cc++) L.A[3 * r + cc] = Jn[3 * r] * R[cc] + Jn[3 * r + 1] * R[3 + cc] + Jn[3 * r + 2] * R[6 + cc];
        L.B[6 * r + 0] = -M[3 * r + 1] * zl + M[3 * r + 2] * yl;
        L.B[6 * r + 1] = M[3 * r + 0] * zl - M[3 * r + 2] * xl;
        L.B[6 * r + 2] = -M[3 * r + 0] * yl + M[3 * r + 1] * xl;
        L.B[6 * r + 3] = M[3 * r + 0];
        L.B[6 * r + 4] = M[3 * r + 1];
        L.B[6 * r + 5] = M[3 * r + 2];
    }
}

// RobustKernelHuber::robustify (Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91)
BA_DEV void huber(double e, double delta, double &rho0, double &rho1) {
    const double dsqr = delta * delta;
    if (e <= dsqr) { rho0 = e; rho1 = 1.0; }
    else {
        const double sqrte = sqrt(e);
        rho0 = 2 * sqrte * delta - dsqr;
        rho1 = delta / sqrte;
    }
}

// inverse of the symmetric 3x3 (Hll + lambda I) given its 6 unique entries h = [00 01 02 11 12 22]
// (block_solver.hpp:389 "Dinv = D->inverse()", Eigen's cofactor formula); output same packing.
BA_DEV void sym3_inverse(const double *h, double *d) {
    const double c00 = h[3] * h[5] - h[4] * h[4];
    const double c01 = h[4] * h[2] - h[1] * h[5];
    const double c02 = h[1] * h[4] - h[3] * h[2];
    const double det = h[0] * c00 + h[1] * c01 + h[2] * c02;
    const double id = 1.0 / det;
    d[0] = c00 * id; d[1] = c01 * id; d[2] = c02 * id;
    d[3] = (h[0] * h[5] - h[2] * h[2]) * id;
    d[4] = (h[1] * h[2] - h[0] * h[4]) * id;
    d[5] = (h[0] * h[3] - h[1] * h[1]) * id;
}

// warp all-reduce (xor butterfly): every lane ends with the same, order-fixed sum
BA_DEV double warp_allsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
BA_DEV double warp_allmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// observation meta word
#define META_KIND(m)   ((m) & 3u)
#define META_CAM(m)    (((m) >> 2) & 0xffu)
#define META_RIG(m)    (((m) >> 10) & 0xffu)
#define META_ROBUST    (1u << 18)
#define META_LEVEL1    (1u << 19)
