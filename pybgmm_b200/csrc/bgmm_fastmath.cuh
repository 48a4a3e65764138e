// bgmm_fastmath.cuh -- short-latency fp64 log / exp for the sweep engine's critical path.
//
// CUDA's libdevice log / exp are accurate to < 1 ulp but are long dependent chains (~1.5 k cycles for a lone warp
// for one log + one exp, measured inside f_eval_lane).  Every weight of the sampler needs one of each, on the
// critical path of every round, so these table-driven versions trade a 3 KB table in global memory (L1 resident)
// for chains of ~12-15 dependent operations.  Accuracy (algorithm checked against 60-digit arithmetic): exp within
// 1 ulp on [-700, 100]; log within 1.2e-16 absolute (2.3e-17 for arguments within 1e-2 of 1), i.e. never more than
// the error the rounding of the argument 1 + G q itself induces -- the floor any evaluation of the reference's
// np.log(1 + ...) (gaussian_components.py:248) is subject to.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace bgmm {
namespace fm {

constexpr int LOG_N = 128;   // log: intervals of the mantissa [1, 2)
constexpr int EXP_N = 64;    // exp: 2^(j/64)
// table layout (doubles): [0, 2*LOG_N): (1/c_i, log c_i) with c_i the centre of interval i; then EXP_N values 2^(j/64)
constexpr int TAB_LEN = 2 * LOG_N + EXP_N;

// host: fill the table with libm (correctly rounded to < 1 ulp; the residual polynomials absorb 1/c_i's rounding
// because r = x * (1/c_i) - 1 is computed exactly relative to the stored 1/c_i and log(c_i) is taken of 1/(stored))
inline void fill_table(double *t) {
    for (int i = 0; i < LOG_N; ++i) {
        const double c = 1.0 + (i + 0.5) / LOG_N;
        const double inv = 1.0 / c;
        t[2 * i] = inv;
        t[2 * i + 1] = -log(inv);   // log of the reciprocal actually used
    }
    for (int j = 0; j < EXP_N; ++j) t[2 * LOG_N + j] = exp2((double)j / EXP_N);
}

// log(x) for normal positive x (the engine's arguments are 1 + G q >= 1 and om in (1/64, 1])
__device__ __forceinline__ double f_log(double x, const double *__restrict__ tab) {
    const long long bits = __double_as_longlong(x);
    const int e = (int)((bits >> 52) & 0x7ff) - 1023;
    const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);  // [1, 2)
    const int i = (int)((bits >> (52 - 7)) & (LOG_N - 1));
    const double inv = tab[2 * i], lc = tab[2 * i + 1];   // plain loads: the table may sit in shared memory
    const double r = fma(m, inv, -1.0);                 // |r| <= 2^-8 (+ rounding of inv), exact up to one rounding
    // log1p(r) = r - r^2/2 + r^3/3 - ... ; |r|^9/9 < 2^-75
    double p = fma(r, -1.0 / 8.0, 1.0 / 7.0);
    p = fma(r, p, -1.0 / 6.0);
    p = fma(r, p, 1.0 / 5.0);
    p = fma(r, p, -1.0 / 4.0);
    p = fma(r, p, 1.0 / 3.0);
    p = fma(r, p, -1.0 / 2.0);
    const double r2 = r * r;
    const double l1p = fma(r2, p, r);
    // e ln2 in two pieces so that the sum keeps ~1 ulp
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double ed = (double)e;
    return fma(ed, LN2_HI, lc) + fma(ed, LN2_LO, l1p);
}

// exp(t) for t in [-745, 709]; returns 0 below the caller's cut-off elsewhere
__device__ __forceinline__ double f_exp(double t, const double *__restrict__ tab) {
    const double INV = 9.23324826168936568e+01;          // 64 / ln 2
    const double C_HI = 1.08304244931787252e-02;         // ln2 / 64, high part (27 trailing zero bits)
    const double C_LO = 2.03070420217029510e-10;         //           low part
    const double kd = rint(t * INV);
    const int k = (int)kd;
    double r = fma(-kd, C_HI, t);
    r = fma(-kd, C_LO, r);                               // |r| <= ln2/128
    // exp(r) - 1 = r + r^2/2 + ... + r^6/720 ; r^7/5040 < 2^-65
    double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    p = fma(r, p, 1.0 / 24.0);
    p = fma(r, p, 1.0 / 6.0);
    p = fma(r, p, 0.5);
    p = fma(r * r, p, r);
    const double s = tab[2 * LOG_N + (k & (EXP_N - 1))];
    const double v = fma(s, p, s);                       // 2^(j/64) * exp(r)
    const int q = k >> 6;                                // floor division: k = 64 q + j
    // scale by 2^q in two steps (q may be as low as -1075: the result can be subnormal)
    const int q1 = q / 2, q2 = q - q1;
    const double s1 = __longlong_as_double((long long)(q1 + 1023) << 52);
    const double s2 = __longlong_as_double((long long)(q2 + 1023) << 52);
    return (v * s1) * s2;
}

}  // namespace fm
}  // namespace bgmm
