/*
 * fmd_math.cuh -- device helpers shared by the kernels of both pipelines: byte -> float
 * conversions, packed f32x2 arithmetic (FADD2 / FMUL2 / FFMA2), the atan2 with C99 Annex F
 * zero/sign behaviour (fmd_kernels.cuh and fmd_quality.cuh include this; it defines no kernel).
 */
#ifndef FMD_MATH_CUH
#define FMD_MATH_CUH

#include <cuda_runtime.h>
#include <stdint.h>

/* ------------------------------------------------------------------ helpers */

/* per-byte (u + 1) mod 256: s + 128 where s = (int8)(u - 127) */
__device__ __forceinline__ uint32_t fmd_bias_bytes(uint32_t w) {
    return ((w & 0x7f7f7f7fu) + 0x01010101u) ^ (w & 0x80808080u);
}

/* byte `i` of a biased word -> exact float in [-128, 127] */
template <int I>
__device__ __forceinline__ float fmd_byte_to_float(uint32_t wb) {
    /* 0x4B0000uu = 8388608 + uu */
    return __uint_as_float(__byte_perm(wb, 0x4B000000u, 0x7540 | I)) - 8388736.0f;
}


/* ---- packed f32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2, one issue slot for two lanes) ---- */
typedef unsigned long long fmd_f2;

__device__ __forceinline__ fmd_f2 fmd_pk(float lo, float hi) {
    fmd_f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void fmd_upk(fmd_f2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ fmd_f2 fmd_fma2(fmd_f2 a, fmd_f2 b, fmd_f2 c) {
    fmd_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ fmd_f2 fmd_mul2(fmd_f2 a, fmd_f2 b) {
    fmd_f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ fmd_f2 fmd_add2(fmd_f2 a, fmd_f2 b) {
    fmd_f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ fmd_f2 fmd_bc(float v) { return fmd_pk(v, v); }
__device__ __forceinline__ fmd_f2 fmd_neg2(fmd_f2 v) {
    float lo, hi;
    fmd_upk(v, lo, hi);
    return fmd_pk(-lo, -hi); /* ptxas folds this into the consumer's operand modifier */
}

/* bytes `I` of two biased words -> the exact pair (s_a, s_b), each in [-128, 127] */
template <int I>
__device__ __forceinline__ fmd_f2 fmd_bytes_to_f2(uint32_t wa, uint32_t wb) {
    return fmd_add2(fmd_pk(__uint_as_float(__byte_perm(wa, 0x4B000000u, 0x7540 | I)),
                           __uint_as_float(__byte_perm(wb, 0x4B000000u, 0x7540 | I))),
                    fmd_bc(-8388736.0f));
}

/* -q3 (normalizeInput, matrix.c:100-118): bytes `I` of two RAW words -> (float)u * (float)(2/255) - 1,
 * product and difference rounded separately like the reference's float multiply and its
 * (exact) double subtraction narrowed back to float */
template <int I>
__device__ __forceinline__ fmd_f2 fmd_bytes_to_unit_f2(uint32_t wa, uint32_t wb) {
    const fmd_f2 u = fmd_add2(fmd_pk(__uint_as_float(__byte_perm(wa, 0x4B000000u, 0x7540 | I)),
                                     __uint_as_float(__byte_perm(wb, 0x4B000000u, 0x7540 | I))),
                              fmd_bc(-8388608.0f));
    return fmd_add2(fmd_mul2(u, fmd_bc((float) (2. / 255.))), fmd_bc(-1.0f));
}

__device__ __forceinline__ float fmd_max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fmd_rcp(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}

/* odd minimax polynomials for atan on [0,1] (Remez, evaluated in f32): max abs error
 * 6 terms 1.75e-6, 7 terms 3.2e-7, 8 terms 1.35e-7 */
#define FMD_ATAN_C6 {9.9997721908e-01f, -3.3262282784e-01f, 1.9354037577e-01f, -1.1642648119e-01f, \
                     5.2647350619e-02f, -1.1719135407e-02f, 0.f, 0.f}
#define FMD_ATAN_C7 {9.9999611155e-01f, -3.3317368053e-01f, 1.9807815551e-01f, -1.3233342042e-01f, \
                     7.9623671389e-02f, -3.3604219717e-02f, 6.8117930109e-03f, 0.f}
#define FMD_ATAN_C8 {9.9999933558e-01f, -3.3329860784e-01f, 1.9946565651e-01f, -1.3908629550e-01f, \
                     9.6421973279e-02f, -5.5912326768e-02f, 2.1862957874e-02f, -4.0545672133e-03f}
/* coefficient i of the AT-term polynomial (compile-time indices only) */
template <int AT>
__device__ __forceinline__ constexpr float fmd_atan_c(int i) {
    constexpr float c6[8] = FMD_ATAN_C6, c7[8] = FMD_ATAN_C7, c8[8] = FMD_ATAN_C8;
    return AT == 6 ? c6[i] : (AT == 7 ? c7[i] : c8[i]);
}

/* quadrant / sign fix-ups shared by the scalar and the packed atan2 */
__device__ __forceinline__ float fmd_atan2_finish(float r, float y, float x) {
    if (fabsf(y) > fabsf(x)) r = 1.57079637f - r;
    if (__float_as_int(x) < 0) r = 3.14159274f - r;
    return __uint_as_float(__float_as_uint(r) | (__float_as_uint(y) & 0x80000000u));
}

/* two atan2 at once: the reductions, reciprocals and fix-ups are per lane, the polynomial is packed */
template <int AT>
__device__ __forceinline__ void fmd_atan2_x2(fmd_f2 im2, fmd_f2 re2, float &out_lo, float &out_hi) {
    float yl, yh, xl, xh;
    fmd_upk(im2, yl, yh);
    fmd_upk(re2, xl, xh);
    const float mnl = fminf(fabsf(xl), fabsf(yl)), mnh = fminf(fabsf(xh), fabsf(yh));
    const float rcl = fmd_rcp(fmd_max3(fabsf(xl), fabsf(yl), 1.0e-30f));
    const float rch = fmd_rcp(fmd_max3(fabsf(xh), fabsf(yh), 1.0e-30f));
    const fmd_f2 t = fmd_mul2(fmd_pk(mnl, mnh), fmd_pk(rcl, rch));
    const fmd_f2 s = fmd_mul2(t, t);
    fmd_f2 p = fmd_fma2(fmd_bc(fmd_atan_c<AT>(AT - 1)), s, fmd_bc(fmd_atan_c<AT>(AT - 2)));
#pragma unroll
    for (int i = AT - 3; i >= 0; --i) p = fmd_fma2(p, s, fmd_bc(fmd_atan_c<AT>(i)));
    float rl, rh;
    fmd_upk(fmd_mul2(p, t), rl, rh);
    out_lo = fmd_atan2_finish(rl, yl, xl);
    out_hi = fmd_atan2_finish(rh, yh, xh);
}

/*
 * atan2 with C99 Annex F zero/sign behaviour for finite inputs (what glibc's
 * atan2f gives the reference): octant reduction, one MUFU.RCP, odd minimax
 * polynomial on [0,1], sign-bit fix-ups.  Scalar form, 8 terms: the exact path.
 */
__device__ __forceinline__ float fmd_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mn = fminf(ax, ay);
    float mx, rc;
    /* max3 with a tiny floor: (0,0) gives t = 0 instead of 0/0 */
    asm("max.f32 %0, %1, %2, %3;" : "=f"(mx) : "f"(ax), "f"(ay), "f"(1.0e-30f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(mx));
    const float t = mn * rc;
    const float s = t * t;
    float p = fmd_atan_c<8>(7);
#pragma unroll
    for (int i = 6; i >= 0; --i) p = fmaf(p, s, fmd_atan_c<8>(i));
    float r = p * t;
    if (ay > ax) r = 1.57079637f - r;
    if (__float_as_int(x) < 0) r = 3.14159274f - r;
    return __uint_as_float(__float_as_uint(r) | (__float_as_uint(y) & 0x80000000u));
}

#endif /* FMD_MATH_CUH */
