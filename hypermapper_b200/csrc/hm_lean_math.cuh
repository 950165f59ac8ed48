// Lean fp64 elementary functions: branch-free, argument ranges known to the callers (the GP covariance transform in
// hm_gp.cu, the EI epilogue in hm_acq.cuh).  The CUDA library sqrt() / exp() / erfc() / division carry inline slow paths
// (denormals, infinities, huge arguments) and cost 40-160 SASS instructions each; these do the same job in 10-30, with
// <= 1-2 ulp error (sqrt, exp, reciprocal) and <= 3e-15 relative (erfcx).  Everything here uses explicit fma(): the
// translation units are compiled with --fmad=false, so nothing else contracts.
#pragma once
#include "hm_common.cuh"

// sqrt(x) for x >= 1e-30 (callers clamp): fp32 rsqrt seed + two coupled Newton steps (Goldschmidt form) + one residual
// correction.
__device__ __forceinline__ double hm_sqrt_pos(double x) {
    float yf;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)x));
    const double y = (double)yf;                     // ~2^-22
    double g = __dmul_rn(x, y);                      // ~ sqrt(x)
    double h = __dmul_rn(0.5, y);                    // ~ 1 / (2 sqrt(x))
    double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);                                // ~2^-43
    e = fma(-g, h, 0.5);
    g = fma(g, e, g);                                // rounding-limited
    const double d = fma(-g, g, x);                  // exact residual
    return fma(d, h, g);
}
// exp(x) for -708 <= x <= 0 (callers clamp): n = rint(x log2 e), f = x - n ln2 (two-constant Cody-Waite), degree-13
// Taylor polynomial on |f| <= ln2/2 (truncation 4e-18), scaled by 2^n through the exponent field (result stays normal).
__device__ __forceinline__ double hm_exp_neg(double x) {
    const double MAGIC = 6755399441055744.0;         // 1.5 * 2^52: the low mantissa word of (t + MAGIC) is rint(t)
    const double t = fma(x, 1.4426950408889634, MAGIC);
    const int n = __double2loint(t);
    const double nf = __dsub_rn(t, MAGIC);
    double f = fma(nf, -6.93147180559945290e-01, x);
    f = fma(nf, -2.31904681384629956e-17, f);
    double p = 1.6059043836821613e-10;               // 1/13!
    p = fma(p, f, 2.08767569878681e-09);             // 1/12!
    p = fma(p, f, 2.505210838544172e-08);            // 1/11!
    p = fma(p, f, 2.755731922398589e-07);            // 1/10!
    p = fma(p, f, 2.7557319223985893e-06);           // 1/9!
    p = fma(p, f, 2.48015873015873e-05);             // 1/8!
    p = fma(p, f, 1.984126984126984e-04);            // 1/7!
    p = fma(p, f, 1.388888888888889e-03);            // 1/6!
    p = fma(p, f, 8.333333333333333e-03);            // 1/5!
    p = fma(p, f, 4.1666666666666664e-02);           // 1/4!
    p = fma(p, f, 1.6666666666666666e-01);           // 1/3!
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// 1 / x for a normal, positive x (no zero, infinity, denormal): MUFU.RCP64H seed (~2^-20) + two Newton steps
__device__ __forceinline__ double hm_rcp_pos(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// sqrt(x) and 1 / sqrt(x) for 1e-30 <= x <= 1e30: the iteration of hm_sqrt_pos, both results kept
__device__ __forceinline__ void hm_sqrt_rsqrt_pos(double x, double* sq, double* rsq) {
    float yf;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)x));
    const double y = (double)yf;
    double g = __dmul_rn(x, y);
    double h = __dmul_rn(0.5, y);
    double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    const double d = fma(-g, g, x);
    *sq = fma(d, h, g);
    *rsq = __dadd_rn(h, h);
}

// erfcx(z) = exp(z^2) erfc(z) for 0 <= z <= 26:  P(s) / (1 + 2 z),  s = c1 z / (z + 4) - 1 in [-1, 1],  P = the degree-20
// Chebyshev interpolant of (1 + 2 z) erfcx(z) in the monomial basis (coefficients and the error table: tools/lean_ei_fit.py;
// max relative error 2e-15 against mpmath).  One reciprocal serves both quotients.
static __constant__ double hm_erfcx_poly[21] = {
    1.25168698288894809e+00,  -1.20437251130418865e-01, -1.89343983902282333e-02, 9.25959156263617550e-02,
    -1.04193434833167717e-01, 7.97132376143787902e-02,  -4.66680821379410402e-02, 2.12474218275156800e-02,
    -7.23771954853323803e-03, 1.59013622279397643e-03,  -6.61037152967160782e-05, -9.57359393404409113e-05,
    3.03439925124943956e-05,  9.68069372030900567e-07,  -2.78641332936506663e-06, 3.53516828912409160e-07,
    2.11267113694037113e-07,  -5.11366939880362601e-08, -1.54579605504629449e-08, 3.86911956016538533e-09,
    8.87799557598591159e-10,
};
#define HM_ERFCX_ZMAX 26.0
__device__ __forceinline__ double hm_erfcx_pos(double z) {
    const double zk = __dadd_rn(z, 4.0);
    const double z2 = fma(2.0, z, 1.0);
    const double r = hm_rcp_pos(__dmul_rn(zk, z2));
    const double r1 = __dmul_rn(r, z2);            // 1 / (z + 4)
    const double r2 = __dmul_rn(r, zk);            // 1 / (1 + 2 z)
    const double s = fma(__dmul_rn(2.3076923076923075, z), r1, -1.0);
    double p = hm_erfcx_poly[20];
#pragma unroll
    for (int k = 19; k >= 0; --k) p = fma(p, s, hm_erfcx_poly[k]);
    return __dmul_rn(p, r2);
}
