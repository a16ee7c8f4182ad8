// gs_math.cuh — arithmetic that must reproduce the reference's host libm bit for bit.
//
// The angle image (reference src/alfa_gs.cpp:366-370) truncates `atan2f(dz, dx) * 100` to an integer,
// so a device atan2f that differs from the host's in the last ulp flips centi-radian bins and, through
// the |angle difference| <= K edge test (Labeler/label.cpp:99), whole ground components.  glibc's
// atan2f/atanf (sysdeps/ieee754/flt-32, the fdlibm algorithm: argument reduction at 7/16, 11/16,
// 19/16, 39/16 + an 11-term odd/even polynomial, all in binary32 without FMA) is therefore restated
// here operation by operation with round-to-nearest intrinsics.  tests/test_math_clone.py compiles
// this header for the host and checks it against the box's libm on 10^7 inputs; on glibc 2.39 it is
// bit-identical (2*10^8 inputs probed).
//
// The same header compiles for the host (GS_MATH_HOST, plain operators under -ffp-contract=off) so
// the clone can be validated without a GPU.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(GS_MATH_HOST)
#define GS_FN __device__ __forceinline__
#define GS_MUL(a, b) __fmul_rn((a), (b))
#define GS_ADD(a, b) __fadd_rn((a), (b))
#define GS_SUB(a, b) __fsub_rn((a), (b))
#define GS_DIV(a, b) __fdiv_rn((a), (b))
GS_FN int32_t gs_f2i(float f) { return __float_as_int(f); }
GS_FN float gs_i2f(int32_t i) { return __int_as_float(i); }
#else
#define GS_FN static inline
#define GS_MUL(a, b) ((a) * (b))
#define GS_ADD(a, b) ((a) + (b))
#define GS_SUB(a, b) ((a) - (b))
#define GS_DIV(a, b) ((a) / (b))
GS_FN int32_t gs_f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
GS_FN float gs_i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
#endif

// atan(0.5), atan(1), atan(1.5), atan(inf) split hi + lo; polynomial coefficients (bit patterns of the
// decimal literals in glibc's s_atanf.c — note aT[0] is 0x3eaaaaab, not the 0x3eaaaaaa of its comment).
#define GS_ATANHI0 0x3eed6338
#define GS_ATANHI1 0x3f490fda
#define GS_ATANHI2 0x3f7b985e
#define GS_ATANHI3 0x3fc90fda
#define GS_ATANLO0 0x31ac3769
#define GS_ATANLO1 0x33222168
#define GS_ATANLO2 0x33140fb4
#define GS_ATANLO3 0x33a22168

GS_FN float gs_atanf(float x) {
    const float aT0 = gs_i2f(0x3eaaaaab), aT1 = gs_i2f((int32_t)0xbe4ccccd), aT2 = gs_i2f(0x3e124925),
                aT3 = gs_i2f((int32_t)0xbde38e38), aT4 = gs_i2f(0x3dba2e6e), aT5 = gs_i2f((int32_t)0xbd9d8795),
                aT6 = gs_i2f(0x3d886b35), aT7 = gs_i2f((int32_t)0xbd6ef16b), aT8 = gs_i2f(0x3d4bda59),
                aT9 = gs_i2f((int32_t)0xbd15a221), aT10 = gs_i2f(0x3c8569d7);
    const int32_t hx = gs_f2i(x);
    const int32_t ix = hx & 0x7fffffff;
    int id;
    float hi = 0.0f, lo = 0.0f;
    if (ix >= 0x4c000000) {  // |x| >= 2^25
        if (ix > 0x7f800000) return GS_ADD(x, x);  // NaN
        const float r = GS_ADD(gs_i2f(GS_ATANHI3), gs_i2f(GS_ATANLO3));
        return hx > 0 ? r : -r;
    }
    if (ix < 0x3ee00000) {             // |x| < 0.4375
        if (ix < 0x31000000) return x;  // |x| < 2^-29
        id = -1;
    } else {
        x = gs_i2f(ix);  // fabsf
        if (ix < 0x3f980000) {      // |x| < 1.1875
            if (ix < 0x3f300000) {  // 7/16 <= |x| < 11/16
                id = 0; hi = gs_i2f(GS_ATANHI0); lo = gs_i2f(GS_ATANLO0);
                x = GS_DIV(GS_SUB(GS_MUL(2.0f, x), 1.0f), GS_ADD(2.0f, x));
            } else {                // 11/16 <= |x| < 19/16
                id = 1; hi = gs_i2f(GS_ATANHI1); lo = gs_i2f(GS_ATANLO1);
                x = GS_DIV(GS_SUB(x, 1.0f), GS_ADD(x, 1.0f));
            }
        } else {
            if (ix < 0x401c0000) {  // |x| < 2.4375
                id = 2; hi = gs_i2f(GS_ATANHI2); lo = gs_i2f(GS_ATANLO2);
                x = GS_DIV(GS_SUB(x, 1.5f), GS_ADD(1.0f, GS_MUL(1.5f, x)));
            } else {                // 2.4375 <= |x| < 2^25
                id = 3; hi = gs_i2f(GS_ATANHI3); lo = gs_i2f(GS_ATANLO3);
                x = GS_DIV(-1.0f, x);
            }
        }
    }
    const float z = GS_MUL(x, x);
    const float w = GS_MUL(z, z);
    // odd and even halves of sum_{i=0..10} aT[i] z^(i+1)
    float s1 = GS_ADD(aT8, GS_MUL(w, aT10));
    s1 = GS_ADD(aT6, GS_MUL(w, s1));
    s1 = GS_ADD(aT4, GS_MUL(w, s1));
    s1 = GS_ADD(aT2, GS_MUL(w, s1));
    s1 = GS_ADD(aT0, GS_MUL(w, s1));
    s1 = GS_MUL(z, s1);
    float s2 = GS_ADD(aT7, GS_MUL(w, aT9));
    s2 = GS_ADD(aT5, GS_MUL(w, s2));
    s2 = GS_ADD(aT3, GS_MUL(w, s2));
    s2 = GS_ADD(aT1, GS_MUL(w, s2));
    s2 = GS_MUL(w, s2);
    const float xs = GS_MUL(x, GS_ADD(s1, s2));
    if (id < 0) return GS_SUB(x, xs);
    const float r = GS_SUB(hi, GS_SUB(GS_SUB(xs, lo), x));
    return hx < 0 ? -r : r;
}

GS_FN float gs_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = gs_i2f(0x3f490fdb), pi_o_2 = gs_i2f(0x3fc90fdb), pi = gs_i2f(0x40490fdb),
                pi_lo = gs_i2f((int32_t)0xb3bbbd2e);
    const int32_t hx = gs_f2i(x), hy = gs_f2i(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return GS_ADD(x, y);  // NaN
    if (hx == 0x3f800000) return gs_atanf(y);                       // x == 1.0
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);              // 2*sign(x) + sign(y)
    if (iy == 0) {                                                  // y == 0
        if (m < 2) return y;
        return m == 2 ? GS_ADD(pi, tiny) : GS_SUB(-pi, tiny);
    }
    if (ix == 0) return hy < 0 ? GS_SUB(-pi_o_2, tiny) : GS_ADD(pi_o_2, tiny);  // x == 0
    if (ix == 0x7f800000) {                                                     // x == inf
        if (iy == 0x7f800000) {
            switch (m) {
                case 0: return GS_ADD(pi_o_4, tiny);
                case 1: return GS_SUB(-pi_o_4, tiny);
                case 2: return GS_ADD(GS_MUL(3.0f, pi_o_4), tiny);
                default: return GS_SUB(GS_MUL(-3.0f, pi_o_4), tiny);
            }
        }
        switch (m) {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return GS_ADD(pi, tiny);
            default: return GS_SUB(-pi, tiny);
        }
    }
    if (iy == 0x7f800000) return hy < 0 ? GS_SUB(-pi_o_2, tiny) : GS_ADD(pi_o_2, tiny);  // y == inf
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = GS_ADD(pi_o_2, GS_MUL(0.5f, pi_lo));  // |y/x| > 2^60
    else if (hx < 0 && k < -60) z = 0.0f;                // |y|/x < -2^60
    else z = gs_atanf(gs_i2f(gs_f2i(GS_DIV(y, x)) & 0x7fffffff));
    switch (m) {
        case 0: return z;
        case 1: return gs_i2f(gs_f2i(z) ^ (int32_t)0x80000000);
        case 2: return GS_SUB(pi, GS_SUB(z, pi_lo));
        default: return GS_SUB(GS_SUB(z, pi_lo), pi);
    }
}
