// Bit-exact restatement of glibc 2.39's atan2f/atanf (sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c,
// the fdlibm single-precision algorithm) for host AND device.
//
// Why: the reference's KannalaBrandt8::project(Eigen::Vector3d) calls the FLOAT functions
// atan2f/sqrtf on double arguments (src/CameraModels/KannalaBrandt8.cpp:47-49). glibc's atan2f is not
// correctly rounded (it differs from (float)atan2((double)y,(double)x) on ~16 % of inputs), so the GPU
// must reproduce the same algorithm, with no FMA contraction, to get the same residuals.
// tests/test_oracle.py::test_fdlibm_atan2f_port_matches_libm checks this file against libm's atan2f on the host
// (hundreds of millions of samples were bit-identical when this was written);
// tests/test_gpu_parity.py::test_device_atan2f_is_libm_exact checks the device build on the GPU.
//
// The algorithm and its constants are fdlibm's (e_atan2f.c / s_atanf.c, "Conversion to float by Ian Lance Taylor,
// Cygnus Support"), whose notice must be preserved:
//
//   ====================================================
//   Copyright (C) 1993 by Sun Microsystems, Inc. All rights reserved.
//
//   Developed at SunPro, a Sun Microsystems, Inc. business.
//   Permission to use, copy, modify, and distribute this
//   software is freely granted, provided that this notice
//   is preserved.
//   ====================================================
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define BAF_HD __host__ __device__ __forceinline__
#define BAF_MUL(a, b) __fmul_rn((a), (b))
#define BAF_ADD(a, b) __fadd_rn((a), (b))
#define BAF_SUB(a, b) __fsub_rn((a), (b))
#define BAF_DIV(a, b) __fdiv_rn((a), (b))
#define BAF_BITS(x) ((int32_t)__float_as_int(x))
#define BAF_FLT(u) __int_as_float((int)(u))
#elif defined(__CUDACC__)
#define BAF_HD __host__ __device__ __forceinline__
#define BAF_MUL(a, b) ((a) * (b))
#define BAF_ADD(a, b) ((a) + (b))
#define BAF_SUB(a, b) ((a) - (b))
#define BAF_DIV(a, b) ((a) / (b))
static inline int32_t baf_bits_host(float x) { int32_t u; memcpy(&u, &x, 4); return u; }
static inline float baf_flt_host(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
#define BAF_BITS(x) baf_bits_host(x)
#define BAF_FLT(u) baf_flt_host((uint32_t)(u))
#else
// plain C/C++ host build: compile with -ffp-contract=off
#define BAF_HD static inline
#define BAF_MUL(a, b) ((a) * (b))
#define BAF_ADD(a, b) ((a) + (b))
#define BAF_SUB(a, b) ((a) - (b))
#define BAF_DIV(a, b) ((a) / (b))
static inline int32_t baf_bits_host(float x) { int32_t u; memcpy(&u, &x, 4); return u; }
static inline float baf_flt_host(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
#define BAF_BITS(x) baf_bits_host(x)
#define BAF_FLT(u) baf_flt_host((uint32_t)(u))
#endif

BAF_HD float baf_atanf(float x) {
    // atanhi/atanlo/aT of s_atanf.c, as bit patterns of the decimal literals
    const float hi0 = BAF_FLT(0x3eed6338u), hi1 = BAF_FLT(0x3f490fdau), hi2 = BAF_FLT(0x3f7b985eu), hi3 = BAF_FLT(0x3fc90fdau);
    const float lo0 = BAF_FLT(0x31ac3769u), lo1 = BAF_FLT(0x33222168u), lo2 = BAF_FLT(0x33140fb4u), lo3 = BAF_FLT(0x33a22168u);
    const float a0 = BAF_FLT(0x3eaaaaabu), a1 = BAF_FLT(0xbe4ccccdu), a2 = BAF_FLT(0x3e124925u), a3 = BAF_FLT(0xbde38e38u),
                a4 = BAF_FLT(0x3dba2e6eu), a5 = BAF_FLT(0xbd9d8795u), a6 = BAF_FLT(0x3d886b35u), a7 = BAF_FLT(0xbd6ef16bu),
                a8 = BAF_FLT(0x3d4bda59u), a9 = BAF_FLT(0xbd15a221u), a10 = BAF_FLT(0x3c8569d7u);
    const int32_t hx = BAF_BITS(x);
    const int32_t ix = hx & 0x7fffffff;
    int id;
    float hi = 0.f, lo = 0.f;
    if (ix >= 0x4c000000) {                       // |x| >= 2^25
        if (ix > 0x7f800000) return BAF_ADD(x, x);  // NaN
        return (hx > 0) ? BAF_ADD(hi3, lo3) : BAF_SUB(-hi3, lo3);
    }
    if (ix < 0x3ee00000) {                        // |x| < 0.4375
        if (ix < 0x31000000) return x;            // |x| < 2^-29
        id = -1;
    } else {
        x = BAF_FLT(ix);                          // fabsf
        if (ix < 0x3f980000) {                    // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; hi = hi0; lo = lo0; x = BAF_DIV(BAF_SUB(BAF_MUL(2.0f, x), 1.0f), BAF_ADD(2.0f, x)); }
            else                 { id = 1; hi = hi1; lo = lo1; x = BAF_DIV(BAF_SUB(x, 1.0f), BAF_ADD(x, 1.0f)); }
        } else {
            if (ix < 0x401c0000) { id = 2; hi = hi2; lo = lo2; x = BAF_DIV(BAF_SUB(x, 1.5f), BAF_ADD(1.0f, BAF_MUL(1.5f, x))); }
            else                 { id = 3; hi = hi3; lo = lo3; x = BAF_DIV(-1.0f, x); }
        }
    }
    const float z = BAF_MUL(x, x);
    const float w = BAF_MUL(z, z);
    float s1 = BAF_ADD(a8, BAF_MUL(w, a10));
    s1 = BAF_ADD(a6, BAF_MUL(w, s1));
    s1 = BAF_ADD(a4, BAF_MUL(w, s1));
    s1 = BAF_ADD(a2, BAF_MUL(w, s1));
    s1 = BAF_ADD(a0, BAF_MUL(w, s1));
    s1 = BAF_MUL(z, s1);
    float s2 = BAF_ADD(a7, BAF_MUL(w, a9));
    s2 = BAF_ADD(a5, BAF_MUL(w, s2));
    s2 = BAF_ADD(a3, BAF_MUL(w, s2));
    s2 = BAF_ADD(a1, BAF_MUL(w, s2));
    s2 = BAF_MUL(w, s2);
    const float xs = BAF_MUL(x, BAF_ADD(s1, s2));
    if (id < 0) return BAF_SUB(x, xs);
    const float r = BAF_SUB(hi, BAF_SUB(BAF_SUB(xs, lo), x));
    return (hx < 0) ? -r : r;
}

BAF_HD float baf_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = BAF_FLT(0x3f490fdbu), pi_o_2 = BAF_FLT(0x3fc90fdbu), pi = BAF_FLT(0x40490fdbu), pi_lo = BAF_FLT(0xb3bbbd2eu);
    const int32_t hx = BAF_BITS(x), hy = BAF_BITS(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return BAF_ADD(x, y);           // NaN
    if (hx == 0x3f800000) return baf_atanf(y);                              // x == 1.0
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);                      // 2*sign(x) + sign(y)
    if (iy == 0) {
        if (m < 2) return y;
        return (m == 2) ? BAF_ADD(pi, tiny) : BAF_SUB(-pi, tiny);
    }
    if (ix == 0) return (hy < 0) ? BAF_SUB(-pi_o_2, tiny) : BAF_ADD(pi_o_2, tiny);
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
                case 0: return BAF_ADD(pi_o_4, tiny);
                case 1: return BAF_SUB(-pi_o_4, tiny);
                case 2: return BAF_ADD(BAF_MUL(3.0f, pi_o_4), tiny);
                default: return BAF_SUB(BAF_MUL(-3.0f, pi_o_4), tiny);
            }
        } else {
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return BAF_ADD(pi, tiny);
                default: return BAF_SUB(-pi, tiny);
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? BAF_SUB(-pi_o_2, tiny) : BAF_ADD(pi_o_2, tiny);
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = BAF_ADD(pi_o_2, BAF_MUL(0.5f, pi_lo));
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = baf_atanf(BAF_FLT(BAF_BITS(BAF_DIV(y, x)) & 0x7fffffff));
    switch (m) {
        case 0: return z;
        case 1: return BAF_FLT((uint32_t)BAF_BITS(z) ^ 0x80000000u);
        case 2: return BAF_SUB(pi, BAF_SUB(z, pi_lo));
        default: return BAF_SUB(BAF_SUB(z, pi_lo), pi);
    }
}
