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
// Structure (v3): both contractions run on the FP64 tensor path (mma.sync m8n8k4 f64 -> DMMA.8x8x4; same peak as DFMA
// on B200 but 8x fewer instructions and 3x fewer operand loads, so the FP64 pipe is the only busy unit).
//   phase A  dot[i][k] = sum_d xs[d][i] * Xs[k][d]  (DMMA)  ->  r^2  ->  Matern/RBF in the accumulator registers ->
//            Kt (shared memory, stored directly in the A-fragment order phase B reads), partial mu.
//   phase B  T[i][m]   = sum_{k<=m} Kx[i][k] * Linv[m][k]   per 32-row panel of L^-1; a warp owns 32 candidates x 32 rows
//            (16 accumulator tiles); ||T||^2 accumulated per candidate.  L^-1 is pre-tiled on the host into per-panel
//            contiguous B-fragment order, streamed by each warp through its own ring of shared-memory stages with
//            cp.async.bulk (TMA 1-D) + mbarrier.  Panels (cost ~ r+1) are dealt boustrophedon to the warp slots so the
//            triangle balances; the all-zero tiles above the diagonal are skipped.
// Fragment order (m8n8k4, lane l, gid = l >> 2, tig = l & 3):  A[i = gid][k = tig],  B[k = tig][n = gid],
//            C[i = gid][n = 2 tig + {0,1}].
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "hm_common.cuh"
#include "hm_lean_math.cuh"

#define HM_GP_THREADS 256
#define HM_GP_WARPS 8
#define HM_GP_BLK 32              // training points per block / rows per L^-1 panel
#define HM_GP_MAX_D 64
#define HM_GP_MAX_NPAD 768
#define HM_GP_MAX_STAGES 4
#define HM_GP_SMEM_LIMIT (227 * 1024)
#define HM_GP_SMEM_HEADER 256     // 8 warps x 4 stages mbarriers
// large-n path (n_pad > HM_GP_MAX_NPAD): 64 x 64 output tiles, 32-deep k stages
#define HM_GPB_T 64
#define HM_GPB_K 32
#define HM_GPB_MAX_NPAD 16384
#define HM_GPB_CHUNK 8192         // candidates per pass (Kx chunk = n_pad x 8192 x 8 B of stream-ordered scratch)

struct hm_gp {
    int n, n_pad, D, kind;
    int G, stage_k, n_stages, aliased;
    int warps;         // warps per CTA: 4 (two CTAs per SM, see hm_gp_plan) or 8
    size_t smem_bytes;
    double sf2, noise, y_mean, y_std, y_std2;
    double* d_XsF;     // [n_pad/8][Dp/4][32]  training inputs / lengthscale in B-fragment order (zero padded)
    double* d_b2;      // [n_pad]     |Xs_j|^2
    double* d_alpha;   // [n_pad]     (zero padded)
    double* d_LpF;     // panels r = 0..R-1, panel r = [(r+1)*8 k4][4 nb][32 lanes] in B-fragment order
    double* d_ls;      // [D]
    // n_pad > HM_GP_MAX_NPAD: the Kx tile no longer fits shared memory; the model is kept row-major and the prediction
    // runs as two plain kernels over candidate chunks (hm_gp_big_*)
    int big;
    double* d_Xs_rm;    // [n_pad][D]   training inputs / lengthscale, row-major (zero padded)
    double* d_Linv_rm;  // [n_pad][n_pad] inverse Cholesky factor, row-major (lower triangular, zero padded)
};

struct HmGpArgs {
    const double* XsF;
    const double* b2;
    const double* alpha;
    const double* LpF;
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
static int hm_gp_plan_w(int n_pad, int D, int warps, size_t limit, int* G_out, int* stage_k_out, int* n_stages_out,
                        int* aliased_out, size_t* smem_out) {
    const int Dp = (D + 3) & ~3;
    for (int G = warps; G >= 1; G >>= 1) {
        const size_t BM = 32 * (size_t)G;
        const size_t fixed = HM_GP_SMEM_HEADER + 8 * ((size_t)n_pad * BM      // Kt
                                                       + BM                     // a2s
                                                       + (size_t)warps * 32     // red
                                                       + (size_t)(n_pad / HM_GP_BLK) * BM);   // mur
        const size_t xs = 8 * (size_t)Dp * BM;
        // ring geometries in order of preference: (k-steps per stage, stages); 16-step stages halve the per-chunk
        // bookkeeping, deeper rings only help when the stage is short.  First with xs and the ring side by side
        // (the ring prefetches during phase A), then aliased (large n).
        static const int geom[][2] = {{16, 3}, {16, 2}, {8, 4}, {8, 3}, {4, 4}, {4, 3}, {8, 2}, {4, 2}};
        for (int aliased = 0; aliased <= 1; ++aliased) {
            for (const auto& gm : geom) {
                const int stage_k = gm[0], ns = gm[1];
                if (!aliased && ns < 3 && stage_k < 16) continue;
                const size_t ring = (size_t)warps * stage_k * 32 * 8 * ns;
                const size_t total = fixed + (aliased ? std::max(xs, ring) : xs + ring);
                if (total <= limit) {
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
    return 0;
}

// Two 4-warp CTAs per SM when the plan fits half of the shared memory (n up to ~300 training points): the two CTAs run
// their phases independently, so one CTA's phase A (sqrt/exp transform: DFMA chains that leave FP64-unit slots idle) can
// run against the other's phase B (DMMA); the phases of ONE CTA cannot overlap because phase B needs the whole Kx tile.
// C2: 4.13 -> 3.94 ms.  (Delaying the second CTA by half a tile to force the interleave changes nothing: DFMA and DMMA
// share one unit, which both phases together keep ~90 % busy once phase A's real DFMA count is included.)  Larger
// models keep one 8-warp CTA per SM.  HM_B200_GP_WARPS=8 forces the latter (measurements).
static int hm_gp_plan(int n_pad, int D, int* warps_out, int* G_out, int* stage_k_out, int* n_stages_out, int* aliased_out,
                      size_t* smem_out) {
    const char* e = getenv("HM_B200_GP_WARPS");
    const int forced = e ? atoi(e) : 0;
    if (forced != 8 && hm_gp_plan_w(n_pad, D, 4, (size_t)113 * 1024, G_out, stage_k_out, n_stages_out, aliased_out, smem_out)) {
        *warps_out = 4;
        return 1;
    }
    *warps_out = 8;
    return hm_gp_plan_w(n_pad, D, 8, HM_GP_SMEM_LIMIT, G_out, stage_k_out, n_stages_out, aliased_out, smem_out);
}

static int hm_gp_create_big(hm_gp* g, int n, int n_pad, int D, const double* Xs, const double* ls, const double* alpha,
                            const double* Linv, double sf2, double noise, double y_mean, double y_std, int kernel_kind,
                            hm_gp_t** out);

extern "C" int hm_gp_create(int n, int D, const double* Xs, const double* ls, const double* alpha, const double* Linv,
                            double sf2, double noise, double y_mean, double y_std, int kernel_kind, hm_gp_t** out) {
    HM_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    HM_REQUIRE(n >= 1 && D >= 1 && D <= HM_GP_MAX_D, "n >= 1, 1 <= D <= 64");
    HM_REQUIRE(Xs && ls && alpha && Linv, "arrays");
    HM_REQUIRE(kernel_kind == 0 || kernel_kind == 1, "kernel_kind");
    int n_pad = (n + HM_GP_BLK - 1) / HM_GP_BLK * HM_GP_BLK;
    const bool big = n_pad > HM_GP_MAX_NPAD || getenv("HM_B200_GP_FORCE_BIG") != nullptr;
    if (big) n_pad = (n + HM_GPB_T - 1) / HM_GPB_T * HM_GPB_T;
    HM_REQUIRE(n_pad <= HM_GPB_MAX_NPAD, "n <= 16384 training points");
    hm_gp* g = new (std::nothrow) hm_gp();
    if (!g) return HM_E_NOMEM;
    memset(g, 0, sizeof(*g));
    if (big) return hm_gp_create_big(g, n, n_pad, D, Xs, ls, alpha, Linv, sf2, noise, y_mean, y_std, kernel_kind, out);
    g->n = n;
    g->n_pad = n_pad;
    g->D = D;
    g->kind = kernel_kind;
    g->sf2 = sf2;
    g->noise = noise;
    g->y_mean = y_mean;
    g->y_std = y_std;
    g->y_std2 = pow(y_std, 2.0);      // normaliser: var * std**2
    if (!hm_gp_plan(n_pad, D, &g->warps, &g->G, &g->stage_k, &g->n_stages, &g->aliased, &g->smem_bytes)) {
        hm_set_error("hm_gp_create: n = %d, D = %d does not fit the shared-memory plan", n, D);
        delete g;
        return HM_E_UNSUPPORTED;
    }
    const int R = n_pad / HM_GP_BLK;
    const int Dp = (D + 3) & ~3, D4 = Dp / 4;
    std::vector<double> hXs((size_t)n_pad * Dp, 0.0), hb2(n_pad, 0.0), ha(n_pad, 0.0);
    std::vector<double> hLp((size_t)512 * R * (R + 1), 0.0);
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        const int kb8 = j / 8, gid = j % 8;
        for (int d = 0; d < D; ++d) {
            // B fragment of the dot product: B[k = d % 4][n = training point % 8]
            hXs[((size_t)kb8 * D4 + d / 4) * 32 + gid * 4 + (d & 3)] = Xs[(size_t)j * D + d];
            s += Xs[(size_t)j * D + d] * Xs[(size_t)j * D + d];
        }
        hb2[j] = s;
        ha[j] = alpha[j];
        // row m = j of L^-1 lives in panel r = j / 32, column block nb = (j % 32) / 8, fragment column gid = j % 8
        const int r = j / HM_GP_BLK, nb = (j % HM_GP_BLK) / 8;
        double* panel = hLp.data() + (size_t)512 * r * (r + 1);
        for (int k = 0; k <= j; ++k)
            panel[((size_t)(k / 4) * 4 + nb) * 32 + gid * 4 + (k & 3)] = Linv[(size_t)j * n + k];
    }
    cudaError_t e;
#define HM_G_TRY(x)                                             \
    if ((e = (x)) != cudaSuccess) {                             \
        hm_set_error("%s: %s", #x, cudaGetErrorString(e));      \
        hm_gp_destroy(g);                                       \
        return (int)e;                                          \
    }
    HM_G_TRY(cudaMalloc((void**)&g->d_XsF, hXs.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_b2, hb2.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_alpha, ha.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_LpF, hLp.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_ls, (size_t)D * 8));
    HM_G_TRY(cudaMemcpy(g->d_XsF, hXs.data(), hXs.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_b2, hb2.data(), hb2.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_alpha, ha.data(), ha.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_LpF, hLp.data(), hLp.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_ls, ls, (size_t)D * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaDeviceSynchronize());      // consumed on non-blocking caller streams (see hm_forest_create)
#undef HM_G_TRY
    *out = g;
    return 0;
}

extern "C" void hm_gp_destroy(hm_gp_t* g) {
    if (!g) return;
    cudaFree(g->d_Xs_rm);
    cudaFree(g->d_Linv_rm);
    cudaFree(g->d_XsF);
    cudaFree(g->d_b2);
    cudaFree(g->d_alpha);
    cudaFree(g->d_LpF);
    cudaFree(g->d_ls);
    delete g;
}

// lean sqrt / exp of the covariance transform: hm_lean_math.cuh (shared with the EI epilogue)

// min(x, cap) for x >= 0 and a cap whose low word is zero: non-negative doubles order like their high words
__device__ __forceinline__ double hm_cap_pos(double x, double cap) {
    return __hiloint2double(min(__double2hiint(x), __double2hiint(cap)), __double2loint(x));
}

// stationary kernel value from the squared scaled distance (GPy Matern52.K_of_r / RBF.K_of_r).  r2 >= 0 and finite:
// candidates with NaN / inf coordinates are patched in the epilogue.
template <int KIND>
__device__ __forceinline__ double hm_gp_k_of_r2(double r2, double sf2) {
    const double SQRT5 = 2.23606797749979;     // np.sqrt(5.)
    if (KIND == 0) {
        // r2 < 1e-30 is treated as r = 1e-15: k'(0) = 0, so the value differs from k(0) by < 1e-29 relative;
        // r is capped at 316 (exp(-sqrt5 * 316) ~ 1e-307: the smallest normal result, 0 for every purpose here)
        const double r = hm_cap_pos(hm_sqrt_pos(r2 > 1e-30 ? r2 : 1e-30), 316.0);
        const double poly = fma(r, fma(r, 5.0 / 3, SQRT5), 1.0);
        return __dmul_rn(__dmul_rn(sf2, poly), hm_exp_neg(__dmul_rn(-SQRT5, r)));
    }
    return __dmul_rn(sf2, hm_exp_neg(__dmul_rn(-0.5, hm_cap_pos(r2, 1408.0))));
}

// D(8x8) += A(8x4) * B(4x8), fp64 tensor path (DMMA.8x8x4)
__device__ __forceinline__ void hm_dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// Panel order of warp slot `slot` (of S): rounds j = 0, 1, ...; panels dealt in descending size, direction alternating.
__device__ __forceinline__ int hm_gp_panel(int j, int slot, int S, int R) {
    const int idx = j * S + ((j & 1) ? (S - 1 - slot) : slot);
    return R - 1 - idx;      // < 0: no panel in this round
}

template <int KIND, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 2 : 1) hm_gp_kernel(const __grid_constant__ HmGpArgs a) {
    extern __shared__ __align__(128) unsigned char gp_smem[];
    const int G = a.G, BM = 32 * G, NBt = 4 * G, S = NW / G, R = a.n_pad / HM_GP_BLK;
    const int D4 = (a.D + 3) >> 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(gp_smem);
    double* KtF = reinterpret_cast<double*>(gp_smem + HM_GP_SMEM_HEADER);  // [n_pad/4][NBt][32]  A fragments of Kx
    double* a2s = KtF + (size_t)a.n_pad * BM;                                // [BM]
    double* red = a2s + BM;                                                  // [S][BM] == [8][32] partial ||T||^2
    double* mur = red + NW * 32;                                    // [R][BM] partial mu
    double* xsF = mur + (size_t)R * BM;                                      // [D4][NBt][32]  A fragments of x / ls
    const int stage_shift = (a.stage_k == 16) ? 4 : (a.stage_k == 8) ? 3 : 2;
    const size_t stage_dbl = (size_t)a.stage_k * 32;
    const uint32_t stage_bytes = (uint32_t)(stage_dbl * 8);
    double* ring_all = a.aliased ? xsF : xsF + (size_t)D4 * 4 * BM;          // [8 warps][n_stages][stage_k/4][4][32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
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
    for (int j = 0; j * S < R; ++j) {
        const int r = hm_gp_panel(j, slot, S, R);
        if (r >= 0) chunks_per_tile += ((r + 1) * HM_GP_BLK) >> stage_shift;
    }
    // ring cursors (warp uniform)
    int c_stage = 0, c_par = 0, p_stage = 0;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i0 = tile * BM;
        // producer state for this tile (lane 0 only): round, chunks left in the current panel, source cursor
        int p_round = -1, p_in_panel = 0, p_left = chunks_per_tile;
        const double* p_src = a.LpF;
        auto produce = [&]() {      // issue the next chunk of this warp's panel sequence into stage p_stage
            if (p_in_panel == 0) {
                const int r = hm_gp_panel(++p_round, slot, S, R);      // valid while p_left > 0
                p_in_panel = ((r + 1) * HM_GP_BLK) >> stage_shift;
                p_src = a.LpF + (size_t)512 * r * (r + 1);
            }
            HM_DEV_ASSERT(p_stage >= 0 && p_stage < NS && p_left > 0 && p_in_panel > 0 && p_round < 2 * R &&
                          p_src + stage_dbl <= a.LpF + (size_t)512 * R * (R + 1));
            hm_mbar_expect_tx(&wbar[p_stage], stage_bytes);
            hm_bulk_g2s(ring + (size_t)p_stage * stage_dbl, p_src, stage_bytes, &wbar[p_stage]);
            p_src += stage_dbl;
            --p_in_panel;
            --p_left;
            if (++p_stage == NS) p_stage = 0;
        };
        if (!a.aliased && lane == 0) {
            for (int s = 0; s < NS && p_left > 0; ++s) produce();
        }

        // ---- scaled candidates (A-fragment order) + squared norms ------------------------------------------
        for (int e = tid; e < D4 * 4 * BM; e += NW * 32) {
            const int d = e / BM, i = e - d * BM;
            double v = 0.0;
            if (d < a.D) {
                const long long gi = min(i0 + i, a.N - 1);
                v = __ldg(a.X + (size_t)d * a.ldx + gi) / __ldg(a.ls + d);
            }
            xsF[((size_t)(d >> 2) * NBt + (i >> 3)) * 32 + (i & 7) * 4 + (d & 3)] = v;
        }
        __syncthreads();
        if (tid < BM) {
            double s = 0.0;
            const double* col = xsF + (size_t)(tid >> 3) * 32 + (tid & 7) * 4;
            for (int d = 0; d < a.D; ++d) {
                const double x = col[(size_t)(d >> 2) * NBt * 32 + (d & 3)];
                s = __dadd_rn(s, __dmul_rn(x, x));
            }
            a2s[tid] = s;
        }
        __syncthreads();

        // ---- phase A: Kx[i][k] = k(x_i, X_k) into KtF, partial mu ------------------------------------------
        // item = (8 candidates, 32 training points): 4 accumulator tiles, the transform runs in registers
        for (int item = warp; item < NBt * R; item += NW) {
            const int iblk = item % NBt, kb = item / NBt;
            double c[4][2];
#pragma unroll
            for (int t = 0; t < 4; ++t) c[t][0] = c[t][1] = 0.0;
            const double* xa = xsF + (size_t)iblk * 32 + lane;
            const double* xb = a.XsF + (size_t)kb * 4 * D4 * 32 + lane;
#pragma unroll 2
            for (int d4 = 0; d4 < D4; ++d4) {
                const double av = xa[(size_t)d4 * NBt * 32];
                double bv[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) bv[t] = __ldg(xb + ((size_t)t * D4 + d4) * 32);
#pragma unroll
                for (int t = 0; t < 4; ++t) hm_dmma(c[t], av, bv[t]);
            }
            // thread holds dot[i = iblk*8 + gid][k = kb*32 + t*8 + 2 tig + e]
            const double a2 = a2s[iblk * 8 + gid];
            double mu = 0.0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int k = kb * HM_GP_BLK + t * 8 + 2 * tig;
                const double2 b2k = __ldg(reinterpret_cast<const double2*>(a.b2 + k));
                const double2 alk = __ldg(reinterpret_cast<const double2*>(a.alpha + k));     // zero in the padding
                // r^2 = -2 a.b + (|b|^2 + |a|^2), clipped at 0 (GPy Stationary._unscaled_dist)
                double r0 = __dadd_rn(__dmul_rn(-2.0, c[t][0]), __dadd_rn(b2k.x, a2));
                double r1 = __dadd_rn(__dmul_rn(-2.0, c[t][1]), __dadd_rn(b2k.y, a2));
                r0 = r0 > 0.0 ? r0 : 0.0;
                r1 = r1 > 0.0 ? r1 : 0.0;
                double v0 = hm_gp_k_of_r2<KIND>(r0, a.sf2);
                double v1 = hm_gp_k_of_r2<KIND>(r1, a.sf2);
                if (k >= a.n) v0 = 0.0;                                 // padding rows contribute nothing
                if (k + 1 >= a.n) v1 = 0.0;
                mu = fma(v0, alk.x, mu);
                mu = fma(v1, alk.y, mu);
                // A-fragment order for phase B: k4 = k / 4, lane' = gid * 4 + k % 4
                double* dst = KtF + ((size_t)((kb * 8 + t * 2 + (tig >> 1)) * NBt + iblk)) * 32 + gid * 4 + 2 * (tig & 1);
                *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
            }
            mu += __shfl_xor_sync(0xffffffffu, mu, 1);
            mu += __shfl_xor_sync(0xffffffffu, mu, 2);
            if (tig == 0) mur[(size_t)kb * BM + iblk * 8 + gid] = mu;
        }
        if (a.aliased) hm_fence_proxy_async();
        __syncthreads();
        if (a.aliased && lane == 0) {
            for (int s = 0; s < NS && p_left > 0; ++s) produce();
        }

        // ---- phase B: T = Kx . Linv^T panel by panel, ||T||^2 per candidate --------------------------------
        double ss[4] = {0.0, 0.0, 0.0, 0.0};
        const double* ka = KtF + (size_t)gB * 4 * 32 + lane;
        const int k4_per_chunk = a.stage_k >> 2;
        for (int j = 0; j * S < R; ++j) {
            const int r = hm_gp_panel(j, slot, S, R);
            if (r < 0) continue;
            double c[4][4][2];
#pragma unroll
            for (int ib = 0; ib < 4; ++ib)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) c[ib][nb][0] = c[ib][nb][1] = 0.0;
            const int nch = ((r + 1) * HM_GP_BLK) >> stage_shift;
            int k4 = 0;
            for (int ch = 0; ch < nch; ++ch) {
                HM_DEV_ASSERT(c_stage >= 0 && c_stage < NS && (c_par == 0 || c_par == 1));
                hm_mbar_wait(&wbar[c_stage], (uint32_t)c_par);
                const double* sb = ring + (size_t)c_stage * stage_dbl + lane;
                for (int q = 0; q < k4_per_chunk; ++q, ++k4) {
                    // tiles strictly above the diagonal of the last block are all zero: rows nb*8.. < k
                    const int k4l = k4 - 8 * r;
                    const int nb0 = (k4l >= 2) ? (k4l >> 1) : 0;
                    double av[4], bv[4];
                    const double* kr = ka + (size_t)k4 * NBt * 32;
#pragma unroll
                    for (int ib = 0; ib < 4; ++ib) av[ib] = kr[ib * 32];
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) bv[nb] = sb[(q * 4 + nb) * 32];
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) {
                        if (nb >= nb0) {
#pragma unroll
                            for (int ib = 0; ib < 4; ++ib) hm_dmma(c[ib][nb], av[ib], bv[nb]);
                        }
                    }
                }
                if (++c_stage == NS) {
                    c_stage = 0;
                    c_par ^= 1;
                }
                __syncwarp();
                // every lane's reads of this stage have completed (their DMMAs were issued above), so the
                // stage can be handed back to the TMA engine without a cross-proxy fence
                if (lane == 0 && p_left > 0) produce();
            }
#pragma unroll
            for (int ib = 0; ib < 4; ++ib)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    ss[ib] = fma(c[ib][nb][0], c[ib][nb][0], ss[ib]);
                    ss[ib] = fma(c[ib][nb][1], c[ib][nb][1], ss[ib]);
                }
        }
#pragma unroll
        for (int ib = 0; ib < 4; ++ib) {
            ss[ib] += __shfl_xor_sync(0xffffffffu, ss[ib], 1);
            ss[ib] += __shfl_xor_sync(0xffffffffu, ss[ib], 2);
        }
        if (tig == 0) {
#pragma unroll
            for (int ib = 0; ib < 4; ++ib) red[slot * BM + gB * 32 + ib * 8 + gid] = ss[ib];
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
            const double a2 = a2s[tid];
            const bool bad = !(a2 < 1.7976931348623157e308);   // NaN / inf coordinates: GPy's K is NaN, so are mean and var
            a.mean[i0 + tid] = bad ? a2 - a2 : mean;
            a.var[i0 + tid] = bad ? a2 - a2 : var;
        }
        __syncthreads();
    }
}

// ============================================================================================================
// Large models (n_pad > 768): the [candidates x n] cross-covariance tile does not fit shared memory any more, so the
// prediction runs over chunks of HM_GPB_CHUNK candidates as
//   hm_gp_big_kx      Kt[k][i] = k(x_i, X_k) (fp64, stream-ordered scratch, k-major so both kernels touch it coalesced),
//                     mu_i = sum_k Kt[k][i] alpha_k (fixed order)
//   hm_gp_big_tri     P[i][rj] = sum_{r in row tile rj} ( sum_{k <= r} Kt[k][i] Linv[r][k] )^2 : a lower-triangular
//                     fp64 GEMM on the FP64 tensor path (DMMA.8x8x4), 64 candidates x 64 rows per CTA, operands staged
//                     through shared memory in 32-deep k slices, the squares reduced in registers
//   hm_gp_big_finish  var_i from sum_rj P[i][rj] (fixed order) with the same epilogue as hm_gp_kernel
// Same arithmetic model as the fused kernel (GPy's GEMM-form distance, clip, lean sqrt / exp); throughput is lower (the
// operands come from L2 instead of a per-warp TMA ring) -- this path exists so that n has no hard limit.
// ============================================================================================================
struct HmGpBigArgs {
    const double* Xs;      // [n_pad][D]
    const double* b2;
    const double* alpha;
    const double* Linv;    // [n_pad][n_pad]
    const double* ls;
    const double* X;       // [D][ldx]
    long long ldx, i0;
    int nc, nc_pad;        // candidates of this chunk, rounded up to 64
    int n, n_pad, D;
    double sf2, noise, y_mean, y_std, y_std2;
    double* Kt;            // [n_pad][nc_pad]
    double* mu;            // [nc_pad]
    double* a2;            // [nc_pad]  |x_i / ls|^2 (NaN / inf marks a bad candidate)
    double* P;             // [nc_pad][n_pad / 64]
    double* mean;
    double* var;
};

template <int KIND>
__global__ void __launch_bounds__(128) hm_gp_big_kx(const HmGpBigArgs a) {
    extern __shared__ double xs_sh[];                       // [D][128] this CTA's candidates / lengthscale
    const int t = threadIdx.x;
    const int i = blockIdx.x * 128 + t;                      // chunk-relative candidate
    const bool live = i < a.nc;
    double a2 = 0.0;
    for (int d = 0; d < a.D; ++d) {
        const double x = live ? a.X[(size_t)d * (size_t)a.ldx + (size_t)(a.i0 + i)] / a.ls[d] : 0.0;
        xs_sh[d * 128 + t] = x;
        a2 = __dadd_rn(a2, __dmul_rn(x, x));
    }
    const bool bad = !(a2 < 1.7976931348623157e308);
    double mu = 0.0;
    for (int k = 0; k < a.n_pad; ++k) {
        double kv = 0.0;
        if (k < a.n && !bad) {
            const double* xk = a.Xs + (size_t)k * a.D;       // the same address for the whole warp: broadcast
            double dot = 0.0;
            for (int d = 0; d < a.D; ++d) dot = fma(xs_sh[d * 128 + t], xk[d], dot);
            double r2 = __dadd_rn(__dmul_rn(-2.0, dot), __dadd_rn(a2, a.b2[k]));
            r2 = r2 > 0.0 ? r2 : 0.0;                        // np.clip(r2, 0, inf)
            kv = hm_gp_k_of_r2<KIND>(r2, a.sf2);
            mu = fma(kv, a.alpha[k], mu);
        }
        if (i < a.nc_pad) a.Kt[(size_t)k * a.nc_pad + i] = kv;
    }
    if (i < a.nc_pad) {
        a.mu[i] = mu;
        a.a2[i] = live ? a2 : 0.0;
    }
}

// CTA (ci, rj): candidates [64 ci, 64 ci + 64) x rows [64 rj, 64 rj + 64); 8 warps, warp w owns candidates 8 w .. 8 w + 7
__global__ void __launch_bounds__(256) hm_gp_big_tri(const HmGpBigArgs a) {
    constexpr int SA = HM_GPB_T + 4, SB = HM_GPB_K + 4;     // padded strides (in doubles): conflict-free fragment loads
    extern __shared__ double big_sh[];
    double (*As)[HM_GPB_K * SA] = reinterpret_cast<double (*)[HM_GPB_K * SA]>(big_sh);                         // [k][candidate]
    double (*Bs)[HM_GPB_T * SB] = reinterpret_cast<double (*)[HM_GPB_T * SB]>(big_sh + 2 * HM_GPB_K * SA);     // [row][k]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int c0 = blockIdx.x * HM_GPB_T, r0 = blockIdx.y * HM_GPB_T;
    const int k_end = r0 + HM_GPB_T;                         // Linv[r][k] = 0 for k > r
    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = 0.0;
    // a thread moves 8 doubles of each operand per stage: A rows k (64 contiguous candidates), B rows r (32 contiguous k)
    auto load = [&](int buf, int k0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = q * 256 + tid;
            const int ka = e >> 6, ca = e & 63;
            As[buf][ka * SA + ca] = a.Kt[(size_t)(k0 + ka) * a.nc_pad + c0 + ca];
            const int rb = e >> 5, kb = e & 31;
            Bs[buf][rb * SB + kb] = a.Linv[(size_t)(r0 + rb) * a.n_pad + k0 + kb];
        }
    };
    load(0, 0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < k_end; k0 += HM_GPB_K) {
        if (k0 + HM_GPB_K < k_end) load(buf ^ 1, k0 + HM_GPB_K);
#pragma unroll
        for (int kk = 0; kk < HM_GPB_K; kk += 4) {
            const double av = As[buf][(kk + tig) * SA + warp * 8 + gid];
#pragma unroll
            for (int t = 0; t < 8; ++t) hm_dmma(acc[t], av, Bs[buf][(t * 8 + gid) * SB + kk + tig]);
        }
        __syncthreads();
        buf ^= 1;
    }
    double ss = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) ss += acc[t][0] * acc[t][0] + acc[t][1] * acc[t][1];
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    if (tig == 0) a.P[(size_t)(c0 + warp * 8 + gid) * (a.n_pad / HM_GPB_T) + blockIdx.y] = ss;
}

__global__ void hm_gp_big_finish(const HmGpBigArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nc) return;
    const int R = a.n_pad / HM_GPB_T;
    double tot = 0.0;
    for (int r = 0; r < R; ++r) tot += a.P[(size_t)i * R + r];
    double v = a.sf2 - tot;
    v = v > 1e-15 ? v : 1e-15;                     // np.clip(var, 1e-15, inf)
    v += a.noise;                                   // Gaussian likelihood
    const double mean = a.mu[i] * a.y_std + a.y_mean;
    double var = v * a.y_std2;
    if (var < 1e-11) var = 1e-11;                   // models.py:1018
    const double a2 = a.a2[i];
    const bool bad = !(a2 < 1.7976931348623157e308);
    a.mean[a.i0 + i] = bad ? a2 - a2 : mean;
    a.var[a.i0 + i] = bad ? a2 - a2 : var;
}

static int hm_gp_create_big(hm_gp* g, int n, int n_pad, int D, const double* Xs, const double* ls, const double* alpha,
                            const double* Linv, double sf2, double noise, double y_mean, double y_std, int kernel_kind,
                            hm_gp_t** out) {
    g->big = 1;
    g->n = n;
    g->n_pad = n_pad;
    g->D = D;
    g->kind = kernel_kind;
    g->sf2 = sf2;
    g->noise = noise;
    g->y_mean = y_mean;
    g->y_std = y_std;
    g->y_std2 = pow(y_std, 2.0);
    std::vector<double> hXs((size_t)n_pad * D, 0.0), hb2(n_pad, 0.0), ha(n_pad, 0.0), hL((size_t)n_pad * n_pad, 0.0);
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int d = 0; d < D; ++d) {
            hXs[(size_t)j * D + d] = Xs[(size_t)j * D + d];
            s += Xs[(size_t)j * D + d] * Xs[(size_t)j * D + d];
        }
        hb2[j] = s;
        ha[j] = alpha[j];
        for (int k = 0; k <= j; ++k) hL[(size_t)j * n_pad + k] = Linv[(size_t)j * n + k];
    }
    cudaError_t e;
#define HM_G_TRY(x)                                             \
    if ((e = (x)) != cudaSuccess) {                             \
        hm_set_error("%s: %s", #x, cudaGetErrorString(e));      \
        hm_gp_destroy(g);                                       \
        return (int)e;                                          \
    }
    HM_G_TRY(cudaMalloc((void**)&g->d_Xs_rm, hXs.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_b2, hb2.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_alpha, ha.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_Linv_rm, hL.size() * 8));
    HM_G_TRY(cudaMalloc((void**)&g->d_ls, (size_t)D * 8));
    HM_G_TRY(cudaMemcpy(g->d_Xs_rm, hXs.data(), hXs.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_b2, hb2.data(), hb2.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_alpha, ha.data(), ha.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_Linv_rm, hL.data(), hL.size() * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaMemcpy(g->d_ls, ls, (size_t)D * 8, cudaMemcpyHostToDevice));
    HM_G_TRY(cudaDeviceSynchronize());
#undef HM_G_TRY
    *out = g;
    return 0;
}

static int hm_gp_mean_var_big(const hm_gp_t* g, const double* Xcm, int64_t ldx, int64_t N, double* mean_out,
                              double* var_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int R = g->n_pad / HM_GPB_T;
    const int chunk = (int)std::min<int64_t>(HM_GPB_CHUNK, (N + HM_GPB_T - 1) / HM_GPB_T * HM_GPB_T);
    const size_t words = (size_t)g->n_pad * chunk + 2 * (size_t)chunk + (size_t)chunk * R;
    const size_t kx_smem = (size_t)g->D * 128 * sizeof(double);
    const size_t tri_smem = sizeof(double) * 2 * (HM_GPB_K * (HM_GPB_T + 4) + HM_GPB_T * (HM_GPB_K + 4));
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gp_big_tri, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tri_smem));
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gp_big_kx<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kx_smem));
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gp_big_kx<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kx_smem));
    double* buf = nullptr;
    HM_CUDA_TRY(cudaMallocAsync((void**)&buf, words * sizeof(double), st));
    HmGpBigArgs a;
    a.Xs = g->d_Xs_rm;
    a.b2 = g->d_b2;
    a.alpha = g->d_alpha;
    a.Linv = g->d_Linv_rm;
    a.ls = g->d_ls;
    a.X = Xcm;
    a.ldx = ldx;
    a.n = g->n;
    a.n_pad = g->n_pad;
    a.D = g->D;
    a.sf2 = g->sf2;
    a.noise = g->noise;
    a.y_mean = g->y_mean;
    a.y_std = g->y_std;
    a.y_std2 = g->y_std2;
    a.Kt = buf;
    a.mu = buf + (size_t)g->n_pad * chunk;
    a.a2 = a.mu + chunk;
    a.P = a.a2 + chunk;
    a.mean = mean_out;
    a.var = var_out;
    for (int64_t i0 = 0; i0 < N; i0 += chunk) {
        a.i0 = i0;
        a.nc = (int)std::min<int64_t>(chunk, N - i0);
        a.nc_pad = (a.nc + HM_GPB_T - 1) / HM_GPB_T * HM_GPB_T;
        const int kx_grid = (a.nc_pad + 127) / 128;
        if (g->kind == 0) hm_gp_big_kx<0><<<kx_grid, 128, kx_smem, st>>>(a);
        else hm_gp_big_kx<1><<<kx_grid, 128, kx_smem, st>>>(a);
        hm_gp_big_tri<<<dim3(a.nc_pad / HM_GPB_T, R), 256, tri_smem, st>>>(a);
        hm_gp_big_finish<<<(a.nc + 255) / 256, 256, 0, st>>>(a);
    }
    const cudaError_t le = cudaGetLastError();
    cudaFreeAsync(buf, st);
    HM_CUDA_TRY(le);
    return 0;
}

extern "C" int hm_gp_mean_var(const hm_gp_t* g, const double* Xcm, int64_t ldx, int64_t N, double* mean_out,
                              double* var_out, void* stream) {
    HM_REQUIRE(g != nullptr, "gp handle");
    HM_REQUIRE(N >= 0 && ldx >= N, "ldx >= N >= 0");
    if (N == 0) return 0;
    HM_REQUIRE(Xcm && mean_out && var_out, "Xcm/mean/var");
    if (g->big) return hm_gp_mean_var_big(g, Xcm, ldx, N, mean_out, var_out, stream);
    HmGpArgs a;
    a.XsF = g->d_XsF;
    a.b2 = g->d_b2;
    a.alpha = g->d_alpha;
    a.LpF = g->d_LpF;
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
    const int ctas_per_sm = g->warps == 4 ? 2 : 1;
    const int grid = (int)std::min<long long>(n_tiles, (long long)hm_num_sms() * ctas_per_sm);
#define HM_GP_GO(K, W)                                                                                                 \
    do {                                                                                                               \
        HM_CUDA_TRY(cudaFuncSetAttribute(hm_gp_kernel<K, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                         (int)g->smem_bytes));                                                         \
        hm_gp_kernel<K, W><<<grid, W * 32, g->smem_bytes, st>>>(a);                                                    \
    } while (0)
    if (g->kind == 0) {
        if (g->warps == 4) HM_GP_GO(0, 4); else HM_GP_GO(0, 8);
    } else {
        if (g->warps == 4) HM_GP_GO(1, 4); else HM_GP_GO(1, 8);
    }
#undef HM_GP_GO
    HM_CUDA_TRY(cudaGetLastError());
    return 0;
}

HM_CHECKED_EXPORT(gp)
