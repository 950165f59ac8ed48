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
// Structure (v2): a SIMT fp64 "GEMM" micro-kernel used twice per candidate tile.  A warp owns a 32 x 32 output tile
// (32 candidates x 32 training points / rows of L^-1); lanes are an 8 x 4 grid, each thread a 4 x 8 register tile
// (32 fp64 accumulators), so one k-step is 6 shared-memory wavefronts for 32 DFMA warp-instructions.
//   phase A  dot[i][k] = sum_d xs[d][i] * Xs[k][d]  ->  r^2  ->  Matern/RBF  ->  Kt[k][i] (shared memory), partial mu
//   phase B  T[i][m]   = sum_{k<=m} Kt[k][i] * Linv[m][k]   per 32-row panel of L^-1;  ||T||^2 accumulated per candidate.
//            L^-1 is pre-tiled on the host into per-panel contiguous [k][32] slabs, streamed by each warp through its
//            own ring of shared-memory stages with cp.async.bulk (TMA 1-D) + mbarrier, so the DFMA stream never waits
//            on L2.  Panels (cost ~ r+1) are dealt boustrophedon to the warp slots so the triangle balances.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "hm_common.cuh"

#define HM_GP_THREADS 256
#define HM_GP_WARPS 8
#define HM_GP_BLK 32              // training points per block / rows per L^-1 panel
#define HM_GP_MAX_D 64
#define HM_GP_MAX_NPAD 768
#define HM_GP_MAX_STAGES 4
#define HM_GP_SMEM_LIMIT (227 * 1024)
#define HM_GP_SMEM_HEADER 256     // 8 warps x 4 stages mbarriers

struct hm_gp {
    int n, n_pad, D, kind;
    int G, stage_k, n_stages, aliased;
    size_t smem_bytes;
    double sf2, noise, y_mean, y_std, y_std2;
    double* d_XsB;     // [R][D][32]  training inputs / lengthscale, blocked by 32 points (zero padded)
    double* d_b2;      // [n_pad]     |Xs_j|^2
    double* d_alpha;   // [n_pad]     (zero padded)
    double* d_Lp;      // panels r = 0..R-1, panel r = [(r+1)*32 k][32 rows]: Lp[k][p] = Linv[32 r + p][k]
    double* d_ls;      // [D]
};

struct HmGpArgs {
    const double* XsB;
    const double* b2;
    const double* alpha;
    const double* Lp;
    const double* ls;
    const double* X;     // [D][N] column-major candidates
    double* mean;
    double* var;
    long long ldx, N;
    int n, n_pad, D, kind;
    int G, stage_k, n_stages, aliased;
    double sf2, noise, y_mean, y_std, y_std2;
};

// shared-memory plan: candidate groups per CTA (G), ring geometry.  Returns 0 if nothing fits.
static int hm_gp_plan(int n_pad, int D, int* G_out, int* stage_k_out, int* n_stages_out, int* aliased_out,
                      size_t* smem_out) {
    for (int G = 8; G >= 1; G >>= 1) {
        const size_t BM = 32 * (size_t)G;
        const size_t fixed = HM_GP_SMEM_HEADER + 8 * ((size_t)n_pad * BM      // Kt
                                                       + BM                     // a2s
                                                       + HM_GP_WARPS * 32       // red
                                                       + (size_t)(n_pad / HM_GP_BLK) * BM);   // mur
        const size_t xs = 8 * (size_t)D * BM;
        // first choice: xs and the ring side by side (the ring can prefetch during phase A), 8 k-steps per stage
        for (int aliased = 0; aliased <= 1; ++aliased) {
            for (int stage_k = 8; stage_k >= 4; stage_k >>= 1) {
                const size_t stage = (size_t)HM_GP_WARPS * stage_k * 32 * 8;
                for (int ns = HM_GP_MAX_STAGES; ns >= (aliased ? 2 : 3); --ns) {
                    const size_t ring = stage * ns;
                    const size_t total = fixed + (aliased ? std::max(xs, ring) : xs + ring);
                    if (total <= HM_GP_SMEM_LIMIT) {
                        *G_out = G;
                        *stage_k_out = stage_k;
                        *n_stages_out = ns;
                        *aliased_out = aliased;
                        *smem_out = total;
                        return 1;
                    }
                }
            }
        }
    }
    return 0;
}

extern "C" int hm_gp_create(int n, int D, const double* Xs, const double* ls, const double* alpha, const double* Linv,
                            double sf2, double noise, double y_mean, double y_std, int kernel_kind, hm_gp_t** out) {
    HM_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    HM_REQUIRE(n >= 1 && D >= 1 && D <= HM_GP_MAX_D, "n >= 1, 1 <= D <= 64");
    HM_REQUIRE(Xs && ls && alpha && Linv, "arrays");
    HM_REQUIRE(kernel_kind == 0 || kernel_kind == 1, "kernel_kind");
    const int n_pad = (n + HM_GP_BLK - 1) / HM_GP_BLK * HM_GP_BLK;
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
    if (!hm_gp_plan(n_pad, D, &g->G, &g->stage_k, &g->n_stages, &g->aliased, &g->smem_bytes)) {
        hm_set_error("hm_gp_create: n = %d, D = %d does not fit the shared-memory plan", n, D);
        delete g;
        return HM_E_UNSUPPORTED;
    }
    const int R = n_pad / HM_GP_BLK;
    std::vector<double> hXs((size_t)n_pad * D, 0.0), hb2(n_pad, 0.0), ha(n_pad, 0.0);
    std::vector<double> hLp((size_t)1024 * R * (R + 1) / 2, 0.0);
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        const int kb = j / HM_GP_BLK, p = j % HM_GP_BLK;
        for (int d = 0; d < D; ++d) {
            hXs[((size_t)kb * D + d) * HM_GP_BLK + p] = Xs[(size_t)j * D + d];
            s += Xs[(size_t)j * D + d] * Xs[(size_t)j * D + d];
        }
        hb2[j] = s;
        ha[j] = alpha[j];
        // row m = j of L^-1 lives in panel r = kb at column p
        double* panel = hLp.data() + (size_t)1024 * kb * (kb + 1) / 2;
        for (int k = 0; k <= j; ++k) panel[(size_t)k * HM_GP_BLK + p] = Linv[(size_t)j * n + k];
    }
    cudaError_t e;
#define HM_G_TRY(x)                                             \
    if ((e = (x)) != cudaSuccess) {                             \
        hm_set_error("%s: %s", #x, cudaGetErrorString(e));      \
        hm_gp_destroy(g);                                       \
        return (int)e;                                          \
    }
    HM_G_TRY(cudaMalloc((void**)&g->d_XsB, hXs.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_b2, hb2.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_alpha, ha.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_Lp, hLp.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_ls, (size_t)D * 8));
    HM_G_TRY(cudaMemcpy(g->d_XsB, hXs.data(), hXs.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_b2, hb2.data(), hb2.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_alpha, ha.data(), ha.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_Lp, hLp.data(), hLp.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_ls, ls, (size_t)D * 8, cudaMemcpyHostToDevice));
#undef HM_G_TRY
    *out = g;
    return 0;
}

extern "C" void hm_gp_destroy(hm_gp_t* g) {
    if (!g) return;
    cudaFree(g->d_XsB);
    cudaFree(g->d_b2);
    cudaFree(g->d_alpha);
    cudaFree(g->d_Lp);
    cudaFree(g->d_ls);
    delete g;
}

// stationary kernel value from the squared scaled distance (GPy Matern52.K_of_r / RBF.K_of_r)
__device__ __forceinline__ double hm_gp_k_of_r2(double r2, int kind, double sf2) {
    const double SQRT5 = 2.23606797749979;     // np.sqrt(5.)
    if (kind == 0) {
        const double r = sqrt(r2);
        const double poly = 1.0 + SQRT5 * r + (5.0 / 3) * (r * r);
        return sf2 * poly * exp(-SQRT5 * r);
    }
    return sf2 * exp(-0.5 * r2);
}

// one k-step of the 32 x 32 warp tile: acc[c][m] += a[c] * b[m]   (4 candidates x 8 columns per thread)
__device__ __forceinline__ void hm_gp_mac(double (&acc)[4][8], const double2& a01, const double2& a23,
                                          const double2 (&b)[4]) {
    const double av[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[c][2 * j] = fma(av[c], b[j].x, acc[c][2 * j]);
            acc[c][2 * j + 1] = fma(av[c], b[j].y, acc[c][2 * j + 1]);
        }
}

// Panel order of warp slot `slot` (of S): rounds j = 0, 1, ...; panels dealt in descending size, direction alternating.
__device__ __forceinline__ int hm_gp_panel(int j, int slot, int S, int R) {
    const int idx = j * S + ((j & 1) ? (S - 1 - slot) : slot);
    return R - 1 - idx;      // < 0: no panel in this round
}

__global__ void __launch_bounds__(HM_GP_THREADS, 1) hm_gp_kernel(const __grid_constant__ HmGpArgs a) {
    extern __shared__ __align__(128) unsigned char gp_smem[];
    const int G = a.G, BM = 32 * G, S = HM_GP_WARPS / G, R = a.n_pad / HM_GP_BLK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(gp_smem);
    double* Kt = reinterpret_cast<double*>(gp_smem + HM_GP_SMEM_HEADER);   // [n_pad][BM]
    double* a2s = Kt + (size_t)a.n_pad * BM;                                 // [BM]
    double* red = a2s + BM;                                                  // [S][BM] == [8][32] partial ||T||^2
    double* mur = red + HM_GP_WARPS * 32;                                    // [R][BM] partial mu
    double* xs = mur + (size_t)R * BM;                                       // [D][BM] scaled candidates
    const size_t stage_dbl = (size_t)a.stage_k * 32;
    double* ring_all = a.aliased ? xs : xs + (size_t)a.D * BM;               // [8 warps][n_stages][stage_k][32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = lane & 7, rg = lane >> 3;
    const long long n_tiles = (a.N + BM - 1) / BM;
    const int NS = a.n_stages;
    uint64_t* wbar = bars + warp * HM_GP_MAX_STAGES;
    double* ring = ring_all + (size_t)warp * NS * stage_dbl;

    if (lane == 0) {
        for (int s = 0; s < NS; ++s) hm_mbar_init(&wbar[s], 1);
        hm_fence_barrier_init();
    }
    __syncthreads();

    // phase-B role of this warp
    const int gB = warp % G, slot = warp / G;
    int chunks_per_tile = 0;
    for (int j = 0;; ++j) {
        const int r = hm_gp_panel(j, slot, S, R);
        if (j * S >= R) break;
        if (r >= 0) chunks_per_tile += (r + 1) * HM_GP_BLK / a.stage_k;
    }
    // ring cursors (warp uniform)
    int c_stage = 0, c_par = 0, p_stage = 0;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i0 = tile * BM;
        // producer state for this tile
        int p_round = 0, p_chunk = 0, p_left = chunks_per_tile;
        auto produce = [&]() {      // lane 0 only: issue the next chunk of this warp's panel sequence
            int r = hm_gp_panel(p_round, slot, S, R);
            while (r < 0 || p_chunk >= (r + 1) * HM_GP_BLK / a.stage_k) {
                ++p_round;
                p_chunk = 0;
                r = hm_gp_panel(p_round, slot, S, R);
            }
            const double* src = a.Lp + (size_t)1024 * r * (r + 1) / 2 + (size_t)p_chunk * stage_dbl;
            const uint32_t bytes = (uint32_t)(stage_dbl * 8);
            hm_mbar_expect_tx(&wbar[p_stage], bytes);
            hm_bulk_g2s(ring + (size_t)p_stage * stage_dbl, src, bytes, &wbar[p_stage]);
            ++p_chunk;
            --p_left;
            if (++p_stage == NS) p_stage = 0;
        };
        if (!a.aliased && lane == 0) {
            for (int s = 0; s < NS && p_left > 0; ++s) produce();
        }

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

        // ---- phase A: Kt[k][i] = k(x_i, X_k), partial mu --------------------------------------------------
        for (int u = warp; u < G * R; u += HM_GP_WARPS) {
            const int g = u % G, kb = u / G;
            double acc[4][8];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int m = 0; m < 8; ++m) acc[c][m] = 0.0;
            const double* xb = a.XsB + (size_t)kb * a.D * HM_GP_BLK + rg * 2;
            const double* xa = xs + g * 32 + cg * 2;
#pragma unroll 2
            for (int d = 0; d < a.D; ++d) {
                const double2 a01 = *reinterpret_cast<const double2*>(xa + (size_t)d * BM);
                const double2 a23 = *reinterpret_cast<const double2*>(xa + (size_t)d * BM + 16);
                double2 b[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = __ldg(reinterpret_cast<const double2*>(xb + d * HM_GP_BLK + j * 8));
                hm_gp_mac(acc, a01, a23, b);
            }
            // r^2 = -2 a.b + (|b|^2 + |a|^2), clipped at 0 (GPy Stationary._unscaled_dist)
            const double2 q01 = *reinterpret_cast<const double2*>(a2s + g * 32 + cg * 2);
            const double2 q23 = *reinterpret_cast<const double2*>(a2s + g * 32 + cg * 2 + 16);
            const double a2v[4] = {q01.x, q01.y, q23.x, q23.y};
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int k = kb * HM_GP_BLK + (m >> 1) * 8 + rg * 2 + (m & 1);
                const double b2k = __ldg(a.b2 + k);
                double r2[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double v = __dadd_rn(__dmul_rn(-2.0, acc[c][m]), __dadd_rn(b2k, a2v[c]));
                    r2[c] = v > 0.0 ? v : 0.0;
                }
                double* row = Kt + (size_t)k * BM + g * 32 + cg * 2;
                *reinterpret_cast<double2*>(row) = make_double2(r2[0], r2[1]);
                *reinterpret_cast<double2*>(row + 16) = make_double2(r2[2], r2[3]);
            }
            __syncwarp();
            // transform pass over the 32 x 32 block this warp just wrote: lane = (candidate pair, k parity)
            {
                const int p = lane & 15, h = lane >> 4;
                double mu0 = 0.0, mu1 = 0.0;
#pragma unroll 1
                for (int t = 0; t < 16; ++t) {
                    const int k = kb * HM_GP_BLK + h + 2 * t;
                    double2* cell = reinterpret_cast<double2*>(Kt + (size_t)k * BM + g * 32 + p * 2);
                    double2 v = *cell;
                    if (k < a.n) {
                        v.x = hm_gp_k_of_r2(v.x, a.kind, a.sf2);
                        v.y = hm_gp_k_of_r2(v.y, a.kind, a.sf2);
                        const double al = __ldg(a.alpha + k);
                        mu0 = fma(v.x, al, mu0);
                        mu1 = fma(v.y, al, mu1);
                    } else {
                        v.x = 0.0;
                        v.y = 0.0;
                    }
                    *cell = v;
                }
                mu0 += __shfl_xor_sync(0xffffffffu, mu0, 16);
                mu1 += __shfl_xor_sync(0xffffffffu, mu1, 16);
                if (h == 0) *reinterpret_cast<double2*>(mur + (size_t)kb * BM + g * 32 + p * 2) = make_double2(mu0, mu1);
            }
        }
        if (a.aliased) hm_fence_proxy_async();
        __syncthreads();
        if (a.aliased && lane == 0) {
            for (int s = 0; s < NS && p_left > 0; ++s) produce();
        }

        // ---- phase B: T = Kx . Linv^T panel by panel, ||T||^2 per candidate --------------------------------
        double ss[4] = {0.0, 0.0, 0.0, 0.0};
        const double* ka = Kt + gB * 32 + cg * 2;
        for (int j = 0; j * S < R; ++j) {
            const int r = hm_gp_panel(j, slot, S, R);
            if (r < 0) continue;
            double acc[4][8];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int m = 0; m < 8; ++m) acc[c][m] = 0.0;
            const int nch = (r + 1) * HM_GP_BLK / a.stage_k;
            for (int ch = 0; ch < nch; ++ch) {
                hm_mbar_wait(&wbar[c_stage], (uint32_t)c_par);
                const double* sb = ring + (size_t)c_stage * stage_dbl + rg * 2;
                const double* kr = ka + (size_t)ch * a.stage_k * BM;
                for (int k4 = 0; k4 < a.stage_k; k4 += 4) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const double* krow = kr + (size_t)(k4 + kk) * BM;
                        const double2 a01 = *reinterpret_cast<const double2*>(krow);
                        const double2 a23 = *reinterpret_cast<const double2*>(krow + 16);
                        double2 b[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            b[q] = *reinterpret_cast<const double2*>(sb + (k4 + kk) * 32 + q * 8);
                        hm_gp_mac(acc, a01, a23, b);
                    }
                }
                if (++c_stage == NS) {
                    c_stage = 0;
                    c_par ^= 1;
                }
                __syncwarp();
                if (lane == 0 && p_left > 0) {
                    hm_fence_proxy_async();
                    produce();
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int m = 0; m < 8; ++m) ss[c] = fma(acc[c][m], acc[c][m], ss[c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            ss[c] += __shfl_xor_sync(0xffffffffu, ss[c], 8);
            ss[c] += __shfl_xor_sync(0xffffffffu, ss[c], 16);
        }
        if (rg == 0) {
            double* rr = red + slot * BM + gB * 32 + cg * 2;
            *reinterpret_cast<double2*>(rr) = make_double2(ss[0], ss[1]);
            *reinterpret_cast<double2*>(rr + 16) = make_double2(ss[2], ss[3]);
        }
        __syncthreads();
        // ---- epilogue -----------------------------------------------------------------------------------
        if (tid < BM && i0 + tid < a.N) {
            double tot = 0.0;
            for (int s = 0; s < S; ++s) tot += red[s * BM + tid];
            double mu = 0.0;
            for (int kb = 0; kb < R; ++kb) mu += mur[(size_t)kb * BM + tid];
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
    a.XsB = g->d_XsB;
    a.b2 = g->d_b2;
    a.alpha = g->d_alpha;
    a.Lp = g->d_Lp;
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
    a.G = g->G;
    a.stage_k = g->stage_k;
    a.n_stages = g->n_stages;
    a.aliased = g->aliased;
    a.sf2 = g->sf2;
    a.noise = g->noise;
    a.y_mean = g->y_mean;
    a.y_std = g->y_std;
    a.y_std2 = g->y_std2;
    cudaStream_t st = (cudaStream_t)stream;
    const int BM = 32 * g->G;
    const long long n_tiles = (N + BM - 1) / BM;
    const int grid = (int)std::min<long long>(n_tiles, hm_num_sms());
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_bytes));
    hm_gp_kernel<<<grid, HM_GP_THREADS, g->smem_bytes, st>>>(a);
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}
