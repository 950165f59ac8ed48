// Scalarised acquisition arithmetic, shared by the fused forest kernel and hm_acq_score.
// Follows random_scalarizations.py:160-506 operation by operation (same association order, no FMA
// contraction: every product/sum goes through __dmul_rn/__dadd_rn).
#pragma once
#include "hm_common.cuh"
#include "hm_lean_math.cuh"

struct HmAcqDev {
    int kind, scal, M, pad;
    double w[HM_MAX_OBJECTIVES];
    double fmin[HM_MAX_OBJECTIVES];
    double beta;
};

__device__ __forceinline__ double hm_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double hm_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double hm_sub(double a, double b) { return __dadd_rn(a, -b); }

// np.maximum / np.minimum propagate NaN
__device__ __forceinline__ double hm_np_max(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    return a > b ? a : b;
}
__device__ __forceinline__ double hm_np_min(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    return a < b ? a : b;
}

// scipy.stats.norm.cdf -> scipy.special.ndtr (Cephes ndtr.c): same branch structure, CUDA erf/erfc.
__device__ __forceinline__ double hm_ndtr(double a) {
    if (a != a) return a;
    const double SQRTH = 7.07106781186547524401E-1;
    double x = hm_mul(a, SQRTH);
    double z = fabs(x);
    double y;
    if (z < SQRTH) {
        y = hm_add(0.5, hm_mul(0.5, erf(x)));
    } else {
        y = hm_mul(0.5, erfc(z));
        if (x > 0) y = hm_sub(1.0, y);
    }
    return y;
}
// scipy.stats.norm.pdf: exp(-x**2/2.0) / sqrt(2*pi)
__device__ __forceinline__ double hm_norm_pdf(double v) {
    const double NORM_PDF_C = 2.5066282746310002;  // np.sqrt(2*np.pi)
    return exp(-hm_mul(v, v) / 2.0) / NORM_PDF_C;
}

// per-objective expected improvement (random_scalarizations.py:432-437), the reference's operations one by one on the
// CUDA library functions: the slow path of hm_ei_term (non-finite or extreme inputs) and the yardstick of its lean path
static __device__ __noinline__ double hm_ei_term_ref(double mean, double var, double f_min) {
    double x_std = sqrt(var);
    double x_mean = hm_sub(1.0, mean);
    double d = hm_sub(x_mean, f_min);
    double v = d / x_std;
    return hm_add(hm_mul(d, hm_ndtr(v)), hm_mul(x_std, hm_norm_pdf(v)));
}

// The same value for ordinary inputs (1e-30 <= var <= 1e30, |d| <= 1e30, |v| <= 36: nothing near underflow, nothing
// non-finite), ~110 fp64 instructions instead of ~330 and no branch inside:
//   x_std, 1/x_std   one coupled Newton iteration (hm_sqrt_rsqrt_pos); v = d / x_std finished with one residual step, so v
//                    is the correctly rounded quotient the reference computes (the far tail is sensitive to the last bit
//                    of v: a relative 1e-16 there moves d*Phi + x_std*phi by ~v^4 * 1e-16)
//   E = exp(-v^2/2)  once, shared: phi(v) = E / sqrt(2 pi),  Phi(-|v|) = E * erfcx(|v| / sqrt 2) / 2
//   erfcx            one degree-20 polynomial on the whole range (hm_erfcx_pos), no erf / erfc branch
// Against 40-digit arithmetic it is closer to the true EI than the reference's own fp64 formula (tools/lean_ei_fit.py);
// the two differ by <= 2e-13 relative for |v| <= 6 and by the reference's own cancellation error (<= 3e-10) in the far
// negative tail -- inside the near-tie window (exact.NEAR_TIE_REL) the host re-scores with the reference's arithmetic.
__device__ __forceinline__ double hm_ei_term(double mean, double var, double f_min) {
    const double d = hm_sub(hm_sub(1.0, mean), f_min);
    if (var >= 1e-30 && var <= 1e30 && fabs(d) <= 1e30) {          // false for NaNs
        double x_std, r_std;
        hm_sqrt_rsqrt_pos(var, &x_std, &r_std);
        double v = __dmul_rn(d, r_std);
        v = fma(fma(-v, x_std, d), r_std, v);
        const double a = fabs(v);
        if (a <= 36.0) {
            const double E = hm_exp_neg(__dmul_rn(__dmul_rn(-0.5, a), a));
            const double tail = __dmul_rn(__dmul_rn(0.5, E), hm_erfcx_pos(__dmul_rn(a, 7.07106781186547524401E-1)));
            const double Phi = v <= 0.0 ? tail : __dadd_rn(1.0, -tail);
            const double phi = __dmul_rn(E, 0.3989422804014327);
            return fma(d, Phi, __dmul_rn(x_std, phi));
        }
    }
    return hm_ei_term_ref(mean, var, f_min);
}

// mean[m], var[m]: per-objective prediction (TS: mean[m] holds the plain prediction).
// feas: feasibility probability (1.0 when no classifier).
__device__ __forceinline__ double hm_acq_value(const HmAcqDev& p, const double* mean, const double* var,
                                               double feas) {
    const int M = p.M;
    if (p.kind == HM_ACQ_EI) {
        if (p.scal == HM_SCAL_LINEAR) {
            double s = 0.0;
            for (int m = 0; m < M; ++m) s = hm_add(s, hm_mul(hm_ei_term(mean[m], var[m], p.fmin[m]), p.w[m]));
            return hm_mul(hm_mul(-1.0, s), feas);
        } else if (p.scal == HM_SCAL_TCHEBYSHEV) {
            double sp = 0.0, tot = 0.0;
            for (int m = 0; m < M; ++m) {
                double sv = hm_mul(p.w[m], hm_ei_term(mean[m], var[m], p.fmin[m]));
                sp = hm_np_max(sv, sp);
                tot = hm_add(tot, sv);
            }
            return hm_mul(hm_mul(-1.0, hm_add(sp, hm_mul(tot, 0.05))), feas);
        } else {
            double sp = __longlong_as_double(0x7ff0000000000000ll);
            for (int m = 0; m < M; ++m) {
                double sv = hm_mul(p.w[m], hm_ei_term(mean[m], var[m], p.fmin[m]));
                sp = hm_np_min(sv, sp);
            }
            return hm_mul(sp, feas);
        }
    } else if (p.kind == HM_ACQ_UCB) {
        if (p.scal == HM_SCAL_LINEAR) {
            double s = 0.0, bf = 0.0;
            for (int m = 0; m < M; ++m) {
                s = hm_add(s, hm_mul(p.w[m], mean[m]));
                bf = hm_add(bf, hm_mul(p.w[m], var[m]));
            }
            s = hm_sub(s, hm_mul(p.beta, sqrt(bf)));
            return hm_mul(s, feas);
        } else if (p.scal == HM_SCAL_TCHEBYSHEV) {
            double sp = 0.0, tot = 0.0;
            for (int m = 0; m < M; ++m) {
                double sv = hm_mul(p.w[m], hm_sub(mean[m], hm_mul(p.beta, sqrt(var[m]))));
                tot = hm_add(tot, sv);
                sp = hm_np_max(sv, sp);
            }
            sp = hm_add(sp, hm_mul(0.05, tot));
            return hm_mul(sp, feas);
        } else {
            double sp = __longlong_as_double(0x7ff0000000000000ll);
            for (int m = 0; m < M; ++m) {
                double sv = hm_mul(p.w[m], hm_sub(mean[m], hm_mul(p.beta, sqrt(var[m]))));
                sp = hm_np_min(sv, sp);
            }
            return -hm_mul(sp, feas);
        }
    } else {  // TS
        if (p.scal == HM_SCAL_LINEAR) {
            double s = 0.0;
            for (int m = 0; m < M; ++m) s = hm_add(s, hm_mul(p.w[m], mean[m]));
            return hm_mul(s, feas);
        } else if (p.scal == HM_SCAL_TCHEBYSHEV) {
            double sp = 0.0, tot = 0.0;
            for (int m = 0; m < M; ++m) {
                double sv = hm_mul(p.w[m], fabs(mean[m]));
                tot = hm_add(tot, sv);
                sp = hm_np_max(sv, sp);
            }
            sp = hm_add(sp, hm_mul(0.05, tot));
            return hm_mul(sp, feas);
        } else {
            double sp = __longlong_as_double(0x7ff0000000000000ll);
            for (int m = 0; m < M; ++m) {
                double sv = hm_mul(p.w[m], fabs(mean[m]));
                sp = hm_np_min(sv, sp);
            }
            return -hm_mul(sp, feas);
        }
    }
}

static inline int hm_acq_to_dev(const hm_acq_params_t* a, HmAcqDev* d) {
    if (a->n_objectives < 1 || a->n_objectives > HM_MAX_OBJECTIVES) return HM_E_INVALID;
    if (a->kind < 0 || a->kind > 2 || a->scalarization < 0 || a->scalarization > 2) return HM_E_INVALID;
    d->kind = a->kind;
    d->scal = a->scalarization;
    d->M = a->n_objectives;
    d->pad = 0;
    for (int i = 0; i < HM_MAX_OBJECTIVES; ++i) {
        d->w[i] = a->weight[i];
        d->fmin[i] = a->f_min[i];
    }
    d->beta = a->beta;
    return 0;
}
