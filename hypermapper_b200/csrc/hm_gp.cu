// K2: GP posterior mean / variance  (replaces GPy GPRegression.predict as called from
// models.compute_gp_prediction_mean_and_uncertainty, models.py:997-1024).
//
//   Kx[j]  = sf2 * (1 + sqrt5 r + 5/3 r^2) * exp(-sqrt5 r),  r = sqrt(clip(|a|^2 + |b_j|^2 - 2 a.b_j, 0))   (Matern-5/2 ARD,
//            GPy Stationary._unscaled_dist form; a = x*/ls, b_j = X_j/ls)              [RBF: sf2 * exp(-r^2/2)]
//   mu     = Kx . alpha                                  (Posterior._raw_predict)
//   var    = clip(sf2 - || L^-1 Kx ||^2, 1e-15) + noise  (dtrtrs + Gaussian likelihood)
//   mean   = mu * y_std + y_mean ; var = var * y_std^2 ; var < 1e-11 -> 1e-11          (normaliser, models.py:1018)
//
// Everything is fp64: var is a cancellation (sf2 - ||.||^2) whose result can be 1e-8 of its terms and alpha can be
// 1e6 when the reference fixes the noise to 0 (models.py:451-453), so tf32 / 3xtf32 cannot meet the 1e-5 relative
// tolerance (DESIGN.md has the measured error); tcgen05 has no f64 kind, so the contraction runs on the FP64 pipe.
//
// Per CTA: a tile of BM candidates.  Phase A builds the cross-covariance tile transposed in shared memory,
// Kt[k][i] (k-major: lanes read consecutive candidates, conflict free).  Phase B is the dense contraction
// T = Kx . L^-T against the cached inverse Cholesky factor, stored k-major (LinvT[k][m] = Linv[m][k]) so that a warp
// reads 16 consecutive rows m as broadcast 16-byte loads; lanes own candidates, registers own 16 rows (x2 candidates
// for n <= 256): 32 fp64 accumulators per thread, 12 load wavefronts per 32 DFMA warp-instructions -> FP64-pipe bound.
// Only k <= m is visited (L^-1 is lower triangular); 16-row slabs are dealt to the 8 warps boustrophedon so the
// triangular work balances.  ||T||^2 is reduced per candidate across warps in shared memory.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "hm_common.cuh"

#define HM_GP_THREADS 256
#define HM_GP_WARPS 8
#define HM_GP_SLAB 16
#define HM_GP_MAX_D 64
#define HM_GP_MAX_NPAD 768

struct hm_gp {
    int n, n_pad, D, kind;
    double sf2, noise, y_mean, y_std, y_std2;
    double* d_Xs;      // [n_pad][D]  training inputs / lengthscale (zero padded)
    double* d_b2;      // [n_pad]     |Xs_j|^2
    double* d_alpha;   // [n_pad]     (zero padded)
    double* d_LinvT;   // [n_pad][n_pad]  LinvT[k][m] = Linv[m][k]  (zero above the diagonal and in the padding)
    double* d_ls;      // [D]
};

struct HmGpArgs {
    const double* Xs;
    const double* b2;
    const double* alpha;
    const double* LinvT;
    const double* ls;
    const double* X;     // [D][N] column-major candidates
    double* mean;
    double* var;
    long long ldx, N;
    int n, n_pad, D, kind;
    double sf2, noise, y_mean, y_std, y_std2;
};

extern "C" int hm_gp_create(int n, int D, const double* Xs, const double* ls, const double* alpha, const double* Linv,
                            double sf2, double noise, double y_mean, double y_std, int kernel_kind, hm_gp_t** out) {
    HM_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    HM_REQUIRE(n >= 1 && D >= 1 && D <= HM_GP_MAX_D, "n >= 1, 1 <= D <= 64");
    HM_REQUIRE(Xs && ls && alpha && Linv, "arrays");
    HM_REQUIRE(kernel_kind == 0 || kernel_kind == 1, "kernel_kind");
    const int n_pad = (n + HM_GP_SLAB - 1) / HM_GP_SLAB * HM_GP_SLAB;
    if (n_pad > HM_GP_MAX_NPAD) {
        hm_set_error("hm_gp_create: n = %d training points exceeds the supported %d", n, HM_GP_MAX_NPAD);
        return HM_E_UNSUPPORTED;
    }
    hm_gp* g = new (std::nothrow) hm_gp();
    if (!g) return HM_E_NOMEM;
    memset(g, 0, sizeof(*g));
    g->n = n;
    g->n_pad = n_pad;
    g->D = D;
    g->kind = kernel_kind;
    g->sf2 = sf2;
    g->noise = noise;
    g->y_mean = y_mean;
    g->y_std = y_std;
    g->y_std2 = pow(y_std, 2.0);      // normaliser: var * std**2
    std::vector<double> hXs((size_t)n_pad * D, 0.0), hb2(n_pad, 0.0), ha(n_pad, 0.0), hLt((size_t)n_pad * n_pad, 0.0);
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int d = 0; d < D; ++d) {
            hXs[(size_t)j * D + d] = Xs[(size_t)j * D + d];
            s += Xs[(size_t)j * D + d] * Xs[(size_t)j * D + d];
        }
        hb2[j] = s;
        ha[j] = alpha[j];
        for (int k = 0; k <= j; ++k) hLt[(size_t)k * n_pad + j] = Linv[(size_t)j * n + k];
    }
    cudaError_t e;
#define HM_G_TRY(x)                                             \
    if ((e = (x)) != cudaSuccess) {                             \
        hm_set_error("%s: %s", #x, cudaGetErrorString(e));      \
        hm_gp_destroy(g);                                       \
        return (int)e;                                          \
    }
    HM_G_TRY(cudaMalloc((void**)&g->d_Xs, hXs.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_b2, hb2.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_alpha, ha.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_LinvT, hLt.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_ls, (size_t)D * 8));
    HM_G_TRY(cudaMemcpy(g->d_Xs, hXs.data(), hXs.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_b2, hb2.data(), hb2.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_alpha, ha.data(), ha.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_LinvT, hLt.data(), hLt.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_ls, ls, (size_t)D * 8, cudaMemcpyHostToDevice));
#undef HM_G_TRY
    *out = g;
    return 0;
}

extern "C" void hm_gp_destroy(hm_gp_t* g) {
    if (!g) return;
    cudaFree(g->d_Xs);
    cudaFree(g->d_b2);
    cudaFree(g->d_alpha);
    cudaFree(g->d_LinvT);
    cudaFree(g->d_ls);
    delete g;
}

// CPL = candidates per lane (2 -> BM = 64, 1 -> BM = 32)
template <int CPL>
__global__ void __launch_bounds__(HM_GP_THREADS, 1) hm_gp_kernel(const HmGpArgs a) {
    constexpr int BM = 32 * CPL;
    extern __shared__ __align__(16) unsigned char gp_smem[];
    double* Kt = reinterpret_cast<double*>(gp_smem);                 // [n_pad][BM]
    double* xs = Kt + (size_t)a.n_pad * BM;                          // [D][BM] scaled candidates
    double* a2s = xs + (size_t)a.D * BM;                             // [BM]
    double* red = a2s + BM;                                          // [HM_GP_WARPS][BM]  partial ||T||^2
    double* mur = red + HM_GP_WARPS * BM;                            // [HM_GP_THREADS / BM][BM] partial mu

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n_tiles = (a.N + BM - 1) / BM;
    const int n_slabs = a.n_pad / HM_GP_SLAB;
    const double SQRT5 = 2.23606797749979;     // np.sqrt(5.)

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i0 = tile * BM;
        // ---- scaled candidates + squared norms ----------------------------------------------------------
        for (int e = tid; e < a.D * BM; e += HM_GP_THREADS) {
            const int d = e / BM, i = e - d * BM;
            const long long gi = min(i0 + i, a.N - 1);
            xs[e] = __ldg(a.X + (size_t)d * a.ldx + gi) / __ldg(a.ls + d);
        }
        __syncthreads();
        if (tid < BM) {
            double s = 0.0;
            for (int d = 0; d < a.D; ++d) s = __dadd_rn(s, __dmul_rn(xs[d * BM + tid], xs[d * BM + tid]));
            a2s[tid] = s;
        }
        __syncthreads();
        // ---- phase A: Kt[k][i], partial mu ---------------------------------------------------------------
        {
            const int i = tid % BM;
            const int kq = tid / BM;                       // 0 .. THREADS/BM - 1
            constexpr int KSTEP = HM_GP_THREADS / BM;
            const double a2 = a2s[i];
            double mu = 0.0;
            for (int k = kq; k < a.n_pad; k += KSTEP) {
                double kv = 0.0;
                if (k < a.n) {
                    const double* xj = a.Xs + (size_t)k * a.D;
                    double dot = 0.0;
                    for (int d = 0; d < a.D; ++d) dot = fma(xs[d * BM + i], __ldg(xj + d), dot);
                    double r2 = __dadd_rn(__dmul_rn(-2.0, dot), __dadd_rn(__ldg(a.b2 + k), a2));
                    r2 = r2 > 0.0 ? r2 : 0.0;              // np.clip(r2, 0, inf)
                    if (a.kind == 0) {
                        const double r = sqrt(r2);
                        const double poly = 1.0 + SQRT5 * r + (5.0 / 3) * (r * r);
                        kv = a.sf2 * poly * exp(-SQRT5 * r);
                    } else {
                        kv = a.sf2 * exp(-0.5 * r2);
                    }
                    mu = fma(kv, __ldg(a.alpha + k), mu);
                }
                Kt[(size_t)k * BM + i] = kv;
            }
            mur[kq * BM + i] = mu;
        }
        __syncthreads();
        // ---- phase B: T = Kx . Linv^T on the lower triangle, ||T||^2 per candidate ------------------------
        double ss[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) ss[c] = 0.0;
        for (int round = 0; round * HM_GP_WARPS < n_slabs; ++round) {
            // boustrophedon deal: rounds alternate direction so heavy (late) and light (early) slabs pair up
            const int slab = round * HM_GP_WARPS + ((round & 1) ? (HM_GP_WARPS - 1 - warp) : warp);
            if (slab >= n_slabs) continue;
            const int m0 = slab * HM_GP_SLAB;
            double acc[CPL][HM_GP_SLAB];
#pragma unroll
            for (int c = 0; c < CPL; ++c)
#pragma unroll
                for (int m = 0; m < HM_GP_SLAB; ++m) acc[c][m] = 0.0;
            const int kend = m0 + HM_GP_SLAB;              // k <= m
            const double* lt = a.LinvT + m0;
#pragma unroll 2
            for (int k = 0; k < kend; ++k) {
                double kx[CPL];
                if (CPL == 2) {
                    const double2 t = *reinterpret_cast<const double2*>(Kt + (size_t)k * BM + 2 * lane);
                    kx[0] = t.x;
                    kx[CPL - 1] = t.y;
                } else {
                    kx[0] = Kt[(size_t)k * BM + lane];
                }
                const double2* lrow = reinterpret_cast<const double2*>(lt + (size_t)k * a.n_pad);
#pragma unroll
                for (int m = 0; m < HM_GP_SLAB / 2; ++m) {
                    const double2 l2 = __ldg(lrow + m);
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        acc[c][2 * m] = fma(kx[c], l2.x, acc[c][2 * m]);
                        acc[c][2 * m + 1] = fma(kx[c], l2.y, acc[c][2 * m + 1]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CPL; ++c)
#pragma unroll
                for (int m = 0; m < HM_GP_SLAB; ++m) ss[c] = fma(acc[c][m], acc[c][m], ss[c]);
        }
#pragma unroll
        for (int c = 0; c < CPL; ++c) red[warp * BM + CPL * lane + c] = ss[c];
        __syncthreads();
        // ---- epilogue -----------------------------------------------------------------------------------
        if (tid < BM && i0 + tid < a.N) {
            double tot = 0.0;
            for (int w = 0; w < HM_GP_WARPS; ++w) tot += red[w * BM + tid];
            double mu = 0.0;
            for (int q = 0; q < HM_GP_THREADS / BM; ++q) mu += mur[q * BM + tid];
            double v = a.sf2 - tot;
            v = v > 1e-15 ? v : 1e-15;                     // np.clip(var, 1e-15, inf)
            v += a.noise;                                   // Gaussian likelihood
            const double mean = mu * a.y_std + a.y_mean;
            double var = v * a.y_std2;
            if (var < 1e-11) var = 1e-11;                   // models.py:1018
            a.mean[i0 + tid] = mean;
            a.var[i0 + tid] = var;
        }
        __syncthreads();
    }
}

extern "C" int hm_gp_mean_var(const hm_gp_t* g, const double* Xcm, int64_t ldx, int64_t N, double* mean_out,
                              double* var_out, void* stream) {
    HM_REQUIRE(g != nullptr, "gp handle");
    HM_REQUIRE(N >= 0 && ldx >= N, "ldx >= N >= 0");
    if (N == 0) return 0;
    HM_REQUIRE(Xcm && mean_out && var_out, "Xcm/mean/var");
    HmGpArgs a;
    a.Xs = g->d_Xs;
    a.b2 = g->d_b2;
    a.alpha = g->d_alpha;
    a.LinvT = g->d_LinvT;
    a.ls = g->d_ls;
    a.X = Xcm;
    a.mean = mean_out;
    a.var = var_out;
    a.ldx = ldx;
    a.N = N;
    a.n = g->n;
    a.n_pad = g->n_pad;
    a.D = g->D;
    a.kind = g->kind;
    a.sf2 = g->sf2;
    a.noise = g->noise;
    a.y_mean = g->y_mean;
    a.y_std = g->y_std;
    a.y_std2 = g->y_std2;
    cudaStream_t st = (cudaStream_t)stream;
    const int cpl = (g->n_pad <= 256) ? 2 : 1;
    const int BM = 32 * cpl;
    const size_t smem = sizeof(double) * ((size_t)g->n_pad * BM + (size_t)g->D * BM + BM + HM_GP_WARPS * BM +
                                          (HM_GP_THREADS / BM) * BM);
    const long long n_tiles = (N + BM - 1) / BM;
    const int grid = (int)std::min<long long>(n_tiles, hm_num_sms());
    if (cpl == 2) {
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_gp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hm_gp_kernel<2><<<grid, HM_GP_THREADS, smem, st>>>(a);
    } else {
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_gp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hm_gp_kernel<1><<<grid, HM_GP_THREADS, smem, st>>>(a);
    }
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}
