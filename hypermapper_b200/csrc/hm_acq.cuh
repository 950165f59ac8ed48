// Scalarised acquisition arithmetic, shared by the fused forest kernel and hm_acq_score.
// Follows random_scalarizations.py:160-506 operation by operation (same association order, no FMA
// contraction: every product/sum goes through __dmul_rn/__dadd_rn).
#pragma once
#include "hm_common.cuh"

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

// per-objective expected improvement (random_scalarizations.py:432-437)
__device__ __forceinline__ double hm_ei_term(double mean, double var, double f_min) {
    double x_std = sqrt(var);
    double x_mean = hm_sub(1.0, mean);
    double d = hm_sub(x_mean, f_min);
    double v = d / x_std;
    return hm_add(hm_mul(d, hm_ndtr(v)), hm_mul(x_std, hm_norm_pdf(v)));
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
