/*
 * hm_b200.h -- C-ABI of the B200-native surrogate-prediction / acquisition-scoring hot path.
 *
 * One shared library (hypermapper_b200/csrc/libhm_b200.so), plain C linkage, plain pointers and
 * sizes, no torch types.  Every `*_dev` / data-path pointer is a DEVICE pointer unless its name ends
 * in `_host`; every launch goes on the caller's stream (`void* stream` == cudaStream_t, 0 = default).
 * Return value: 0 on success, otherwise a CUDA error code (cudaError_t) or a negative HM_E* code;
 * hm_last_error() returns a static description.  No entry point allocates on the data path except
 * into the workspace that is handed in (handles own their model memory).
 *
 * Each entry point cites the reference (luinardi/hypermapper v2.2.12) interface it replaces.
 */
#ifndef HM_B200_H
#define HM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HM_MAX_OBJECTIVES 8

#define HM_E_INVALID   (-1)   /* bad argument */
#define HM_E_NOMEM     (-2)
#define HM_E_UNSUPPORTED (-3)

/* acquisition kinds: random_scalarizations.py:109-150 */
#define HM_ACQ_EI  0   /* EI   random_scalarizations.py:356-506 */
#define HM_ACQ_UCB 1   /* ucb  random_scalarizations.py:160-261 */
#define HM_ACQ_TS  2   /* thompson_sampling random_scalarizations.py:264-353 */
/* scalarization methods (schema.json "scalarization_method") */
#define HM_SCAL_LINEAR 0
#define HM_SCAL_TCHEBYSHEV 1
#define HM_SCAL_MODIFIED_TCHEBYSHEV 2

const char* hm_last_error(void);
/* library / device probe: returns the SM count of the current device (148 on B200) or <0 */
int hm_device_sm_count(void);
int hm_version(void);
/* burst FP64 DFMA rate of the current device in TFLOP/s (roofline denominator for K2; not on a data path) */
double hm_fp64_peak_tflops(void);
/* -1 in a normal build.  In the checked build (make -C hypermapper_b200/csrc checked; HM_B200_LIB selects it) every index
 * into a hand-rolled shared-memory ring / queue / candidate buffer and every staged-node address is verified on the
 * device; this returns how many of those invariants were violated since the library was loaded (0 = clean) and where
 * (*tu_out: 0 forest, 1 top-k, 2 GP, 3 Pareto, 4 space; *line_out: source line of the last one). */
int hm_checked_failures(int* tu_out, int* line_out);

/* ------------------------------------------------------------------------------------------
 * K1  forest traversal + per-leaf table reduction
 * replaces: RFModel.get_leaves_per_sample (models.py:174-182 -> sklearn Tree._apply_dense),
 *           RFModel.compute_rf_prediction (models.py:225-240),
 *           RFModel.compute_rf_prediction_variance (models.py:242-275),
 *           RandomForestRegressor.predict as used by thompson sampling (models.py:879-890).
 * ------------------------------------------------------------------------------------------ */
typedef struct hm_forest hm_forest_t;

/* Build a device-resident forest from HOST arrays in sklearn layout (all trees concatenated).
 *   tree_offsets[T+1] : node offset of each tree
 *   feature[n]        : split feature, <0 at leaves            (tree_.feature)
 *   threshold[n]      : fp64 split threshold                   (tree_.threshold, after fit_RF re-randomisation)
 *   left/right[n]     : tree-local child ids, -1 at leaves     (tree_.children_left/right)
 *   q/s/u[n]          : per-node fp64 tables, read at leaves only, computed by the host EXACTLY as the
 *                       reference does:  q = m/T  (models.py:238),  s = (m**2)/T (models.py:263-264),
 *                       u = v/T (models.py:260) with m,v the Hutter per-leaf mean / variance
 *                       (models.py:79-122).  May be NULL (then mean/var outputs are unavailable).
 *   value[n]          : tree_.value[:,0,0] (plain sklearn prediction, used by the TS path). May be NULL.
 * The split test is sklearn's `(float32)x <= (float64)thr`; thresholds are stored as the largest
 * float32 <= thr, which gives the identical decision for every finite float32 x.                  */
int hm_forest_create(int n_trees, int n_features, const int64_t* tree_offsets_host,
                     const int32_t* feature_host, const double* threshold_host,
                     const int32_t* left_host, const int32_t* right_host,
                     const double* q_host, const double* s_host, const double* u_host,
                     const double* value_host, hm_forest_t** out);
void hm_forest_destroy(hm_forest_t* f);
int hm_forest_n_trees(const hm_forest_t* f);
int hm_forest_n_features(const hm_forest_t* f);
/* 1 if the forest is streamed through shared memory, 0 if a tree was too large and it is walked in L2 */
int hm_forest_staged(const hm_forest_t* f);
/* number of shared-memory tree groups of the device layout (<= 2 summed over the forests of a call: the forests
 * stay resident in shared memory; more: they stream through it and large pools take the coherence pre-pass) */
int hm_forest_n_groups(const hm_forest_t* f);
/* 1 if the groups of the fp32 layout (packed == 0) / of the packed layout (packed != 0) carry their trees' leaf records:
 * a call whose forests all do and stay resident in shared memory (<= 2 groups in total) never gathers a leaf record from
 * global memory (models.py:225-275's per-leaf dictionaries live next to the nodes) */
int hm_forest_leaf_resident(const hm_forest_t* f, int packed);

/* acquisition parameters, one struct for all kinds (unused fields ignored) */
typedef struct hm_acq_params {
    int32_t kind;            /* HM_ACQ_*  */
    int32_t scalarization;   /* HM_SCAL_* */
    int32_t n_objectives;    /* M <= HM_MAX_OBJECTIVES */
    int32_t reserved;
    double weight[HM_MAX_OBJECTIVES]; /* objective_weights, or reciprocate_weights(...) for modified_tchebyshev
                                         (utility_functions.py:666-681) -- the host passes the ones the branch uses */
    double f_min[HM_MAX_OBJECTIVES];  /* EI: 1 - (min(data[obj]) - lo)/diff  (random_scalarizations.py:429-431) */
    double beta;                       /* UCB: sqrt(0.125*log(2*iter+1))  (random_scalarizations.py:187) */
} hm_acq_params_t;

/* Candidates are a column-major float32 matrix Xcm[f*ldx + i], f < n_features, i < N (ldx >= N):
 * the encoding of models.preprocess_data_array (models.py:941-984) cast to float32 as sklearn does.
 *
 * hm_rf_eval walks M forests (objectives) over the same candidates in ONE launch and emits whichever
 * outputs are non-NULL:
 *   leaf_out[m]  : int32 [T_m][N]  sklearn leaf node id per (tree, candidate)         (A3)
 *   mean_out     : fp64 [M][N]     Hutter mean                                         (A4)
 *   var_out      : fp64 [M][N]     Hutter variance (0 -> 1e-5)                         (A5)
 *   pred_out     : fp64 [M][N]     plain RandomForestRegressor.predict                 (A12)
 *   score_out    : fp64 [N]        scalarised acquisition value (requires acq != NULL) (A10-A12)
 *   feas         : fp64 [N] or NULL  feasibility probability multiplied into the score
 */
/* Coherence pre-pass of hm_rf_eval (candidates bucketed by the leaves of two key trees so that the lanes of a warp
 * walk the same nodes), a property of the handle (the first forest of a call decides): -1 = automatic (forests that
 * stream through shared memory, N >= 2^17; the default of a new handle, which the environment variable
 * HM_B200_FOREST_SORT overrides at creation), 0 = never, 1 = always.  Results are bit-identical either way; this only
 * exists for tests and measurements. */
int hm_forest_set_coherence(hm_forest_t* f, int mode);
int hm_rf_eval(const hm_forest_t* const* forests, int M, const float* Xcm, int64_t ldx, int64_t N,
               int32_t* const* leaf_out, double* mean_out, double* var_out, double* pred_out,
               const hm_acq_params_t* acq, const double* feas, double* score_out, void* stream);

/* Packed candidates (all-discrete search spaces: ordinal / categorical parameters only).  A candidate is ONE 64-bit code
 * -- the rank of every ordinal value and a one-hot field per categorical, parameters in input order from bit 63 downwards
 * (hm_space_packed_field gives every field; csrc/hm_pack.cuh states the layout) -- instead of D_enc float32 columns:
 * 8 bytes per candidate in HBM and over PCIe, and no per-visit feature load in the traversal (the split becomes a shift
 * and an unsigned compare on the code, exact because the encoding of models.py:957-984 is monotone in the rank).
 *   hm_forest_attach_packing  builds the packed node arrays of a forest for the space its model was fitted on (once)
 *   hm_rf_eval_packed         hm_rf_eval on codes; every output is bit-identical to hm_rf_eval on the encoded form of the
 *                             same candidates (sklearn's decision (double)x32 <= threshold, models.py:174-182) */
typedef struct hm_space hm_space_t;
int hm_forest_attach_packing(hm_forest_t* f, const hm_space_t* space);
int hm_forest_packed(const hm_forest_t* f);
int hm_forest_packed_groups(const hm_forest_t* f);
int hm_rf_eval_packed(const hm_forest_t* const* forests, int M, const uint64_t* codes, int64_t N,
                      int32_t* const* leaf_out, double* mean_out, double* var_out, double* pred_out,
                      const hm_acq_params_t* acq, const double* feas, double* score_out, void* stream);

/* Fused scoring + selection: the k best candidates of a pool (the stable (value, index) order of hm_topk =
 * utility_functions.get_min_configurations, utility_functions.py:295-316) WITHOUT the score array ever being written:
 * what local_search needs from its random pools (the seeds of the hill climb, local_search.py:625-645).
 * The forest kernel's epilogue appends a candidate to a short list only if it can still be among the k best (its
 * score beats the k-th best of a 65 536-candidate prefix, or ties it from inside the prefix, or is NaN); hm_topk_merge
 * on that list is the top-k of the pool.  Identical to hm_rf_eval + hm_topk for every input (a list overflow falls
 * back to exactly that).  out_vals[k], out_idx[k]: device; workspace: hm_rf_score_topk_workspace_bytes(k) device bytes.
 * The call synchronises the stream once (the list length decides about the fall-back). */
int64_t hm_rf_score_topk_workspace_bytes(int k);
int hm_rf_score_topk(const hm_forest_t* const* forests, int M, const float* Xcm, int64_t ldx, int64_t N,
                     const hm_acq_params_t* acq, const double* feas, int k, double* out_vals, int64_t* out_idx,
                     void* workspace, void* stream);
int hm_rf_score_topk_packed(const hm_forest_t* const* forests, int M, const uint64_t* codes, int64_t N,
                            const hm_acq_params_t* acq, const double* feas, int k, double* out_vals,
                            int64_t* out_idx, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3  scalarised acquisition on precomputed mean/variance (GP path, or RF mean/var API users)
 * replaces: EI / ucb / thompson_sampling arithmetic (random_scalarizations.py:160-506)
 *   mean, var : fp64 [M][N] (ld = ldm); for HM_ACQ_TS `mean` holds the posterior sample / prediction.
 * ------------------------------------------------------------------------------------------ */
int hm_acq_score(const hm_acq_params_t* acq, const double* mean, const double* var, int64_t ldm,
                 const double* feas, int64_t N, double* score_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * stable k-smallest: replaces utility_functions.get_min_configurations (utility_functions.py:295-316)
 * i.e. k x np.argmin-and-delete  ==  the k lexicographically smallest (value, index) pairs, ascending,
 * NaN ordered first (np.argmin returns the first NaN), -0.0 == +0.0.
 *   out_vals[k], out_idx[k] (device); out_idx = index_base + position.  If N < k the tail is filled
 *   with (+inf, -1).  workspace: hm_topk_workspace_bytes(k) bytes of device memory.
 * ------------------------------------------------------------------------------------------ */
int64_t hm_topk_workspace_bytes(int k);
int hm_topk(const double* vals, int64_t N, int64_t index_base, int k, double* out_vals, int64_t* out_idx,
            void* workspace, void* stream);
/* merge P sorted-or-unsorted candidate lists of (value, global index) pairs (e.g. the all-gathered
 * per-rank top-k) into the global top-k with the same ordering.  n = total pairs. */
int hm_topk_merge(const double* vals, const int64_t* idx, int64_t n, int k, double* out_vals,
                  int64_t* out_idx, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  GP posterior mean / variance
 * replaces: GPy GPRegression.predict as called by compute_gp_prediction_mean_and_uncertainty
 *           (models.py:997-1024): Matern52 ARD cross-covariance, mean = Kx^T alpha,
 *           var = sf2 - ||L^-1 Kx||^2 (clip 1e-15) + noise, de-normalise, floor 1e-11.
 *   Xs_host    : fp64 [n][D] training inputs ALREADY divided by the ARD lengthscales
 *   ls_host    : fp64 [D]    ARD lengthscales (candidates are divided by them on the device, as GPy does)
 *   alpha_host : fp64 [n]    K_y^-1 y_normalised
 *   Linv_host  : fp64 [n][n] row-major inverse of the lower Cholesky factor of K_y (lower triangular)
 *   kernel_kind: 0 = Matern-5/2, 1 = RBF;  D <= 64.  Up to 768 training points the prediction is ONE fused kernel (Kx tile
 *   in shared memory, L^-1 panels through per-warp TMA rings); larger models (n <= 16384) run as a chunked Kx kernel +
 *   a lower-triangular fp64 tensor-path GEMM through stream-ordered scratch (same arithmetic, lower throughput)
 * ------------------------------------------------------------------------------------------ */
typedef struct hm_gp hm_gp_t;
int hm_gp_create(int n, int D, const double* Xs_host, const double* ls_host, const double* alpha_host,
                 const double* Linv_host, double sf2, double noise, double y_mean, double y_std,
                 int kernel_kind, hm_gp_t** out);
void hm_gp_destroy(hm_gp_t* g);
/* Xcm: fp64 column-major candidates [D][N] (ld = ldx), raw (not divided by lengthscale).
 * mean_out/var_out: fp64 [N]; var floored at 1e-11 as models.py:1018.                       */
int hm_gp_mean_var(const hm_gp_t* g, const double* Xcm, int64_t ldx, int64_t N, double* mean_out,
                   double* var_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4  Pareto dominance filter
 * replaces: compute_pareto.sequential_is_pareto_efficient (compute_pareto.py:89-117), both passes.
 *   costs : fp64 [N][M] row-major (device), M <= HM_MAX_OBJECTIVES
 *   mask  : uint8 [N] (device) 1 = Pareto efficient
 *   workspace: hm_pareto_workspace_bytes(N, M) bytes of device memory
 *   n_front_out_host: optional host int64 receiving the number of efficient points
 * The call synchronises the stream internally (the sweep is data dependent).
 * ------------------------------------------------------------------------------------------ */
int64_t hm_pareto_workspace_bytes(int64_t N, int M);
int hm_pareto_filter(const double* costs, int64_t N, int M, uint8_t* mask, void* workspace,
                     int64_t* n_front_out_host, void* stream);
/* pass 1 only (the order-independent weak front, compute_pareto.py:100-104): used by the sharded
 * path, which exchanges pass-1 survivors between ranks before the sequential pass 2. */
int hm_pareto_pass1(const double* costs, int64_t N, int M, uint8_t* mask, void* workspace,
                    int64_t* n_alive_out_host, void* stream);

/* pass 1 + ordered compaction of its survivors, for the sharded path: rows_out[j][M] / idx_out[j] = the j-th surviving
 * row (ascending index; idx = index_base + row) for j < min(*n_alive_out_host, cap), all on the device and ordered on
 * `stream`; *saw_nan_out_host = 1 when the block holds NaN costs (the filter detects them itself). */
int hm_pareto_pass1_compact(const double* costs, int64_t N, int M, uint8_t* mask, void* workspace, double* rows_out,
                            int64_t* idx_out, int64_t cap, int64_t index_base, int64_t* n_alive_out_host,
                            int32_t* saw_nan_out_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5 / K6 / K7  search-space side of the path on the device
 * A space is the list of input parameters in input-parameter order (space.py Real/Integer/Ordinal/
 * CategoricalParameter) plus one fp64 table blob the descriptors point into.
 * ------------------------------------------------------------------------------------------ */
#define HM_P_REAL 0
#define HM_P_INTEGER 1
#define HM_P_ORDINAL 2
#define HM_P_CATEGORICAL 3
/* how get_x_probability / randomly_select of the parameter work */
#define HM_PRIOR_BETA 0      /* real, integer: stats.beta.pdf((x-lo)/(hi-lo), a, b) (space.py:279-286, 405-410); the
                                 "uniform" prior is a = b = 1 */
#define HM_PRIOR_TABLE 1     /* probability per value index: categorical, "distribution" priors, and ordinals with a
                                 named density (cdf differences, space.py:524-533, precomputed by the host) */
#define HM_PRIOR_GAUSSIAN 2  /* real "custom_gaussian": norm.pdf(x, a = mean, b = std) (space.py:272-277) */
#define HM_PRIOR_INTERP 3    /* real list prior: np.interp on the normalised list (space.py:222-229) */
typedef struct hm_param_desc {
    int32_t kind;        /* HM_P_* */
    int32_t n_values;    /* ordinal: number of values; categorical: number of categories */
    int32_t out_col;     /* first encoded column (models.py:957-984: non-categoricals in input order, then the
                            one-hot blocks) */
    int32_t normalize;   /* real/integer: z-score with mean/std (models.py:966-971) */
    int32_t prior_kind;  /* HM_PRIOR_* */
    int32_t n_prior;     /* TABLE: entries; INTERP: length of the prior list */
    int32_t ordinal_beta;/* ordinal / integer with a named density: sampled as Beta(a, b) -> rounded index
                            (space.py:383-387, 496-499) even though its probability is a TABLE */
    int32_t n_cdf;       /* INTERP: length of the sampling tables (10000 in the reference, space.py:127-130) */
    int64_t values_off;  /* ordinal: tables[values_off ..] = n_values ascending values, then n_values encoded
                            values index/size (models.py:962-965, host arithmetic) */
    int64_t prior_off;   /* TABLE: n_prior probabilities, then n_prior cumulative sums (np.random.choice);
                            INTERP: n_prior abscissae linspace(lo, hi, n_prior), then the n_prior densities */
    int64_t cdf_off;     /* INTERP: n_cdf cumulative values, then the n_cdf abscissae (space.py:151-153) */
    double lo, hi;       /* get_min() / get_max() */
    double mean, std;    /* z-score statistics */
    double a, b;         /* Beta alpha/beta, or Gaussian mean/std */
    double lbeta;        /* log B(a, b) */
} hm_param_desc_t;
int hm_space_create(int n_params, const hm_param_desc_t* params_host, const double* tables_host, int64_t n_tables,
                    hm_space_t** out);
void hm_space_destroy(hm_space_t* s);
int hm_space_n_params(const hm_space_t* s);
int hm_space_n_encoded(const hm_space_t* s);
/* packed codes: number of bits a candidate of this space takes (0: the space has real / integer parameters or needs
 * more than 64 bits and cannot be packed); bit range [low_bit, low_bit + width) of one parameter's field (ordinal: the
 * rank of the value; categorical: one-hot, category c = bit low_bit + c) */
int hm_space_packed_bits(const hm_space_t* s);
int hm_space_packed_field(const hm_space_t* s, int param, int* low_bit, int* width);

/* K5  replaces models.preprocess_data_array (models.py:941-984), bit-exact.
 *   raw: fp64 raw parameter values, row-major [N][ldr >= D] (raw_col_major = 0, the layout of the reference's
 *        sample arrays) or column-major [D][ldr >= N]; categoricals as integer indices, ordinals as values.
 *   X32cm / X64cm (either may be NULL): encoded column-major [D_enc][ld].
 *   n_invalid_dev (nullable, caller zeroes it): incremented when an ordinal value is not in its list or a
 *        categorical index is out of range (the reference raises ValueError there); such entries encode as NaN. */
int hm_encode(const hm_space_t* s, const double* raw, int raw_col_major, int64_t ldr, int64_t N, float* X32cm,
              int64_t ld32, double* X64cm, int64_t ld64, int32_t* n_invalid_dev, void* stream);

/* K5 / K6 for packed codes: hm_encode / hm_sample_pool writing one 64-bit code per candidate instead of the encoded
 * columns (same value search, same Philox stream: candidate i of a pool is the same configuration either way). */
int hm_encode_packed(const hm_space_t* s, const double* raw, int raw_col_major, int64_t ldr, int64_t N,
                     uint64_t* codes_out, int32_t* n_invalid_dev, void* stream);
int hm_sample_pool_packed(const hm_space_t* s, uint64_t seed, int64_t index_base, int64_t N, int use_priors,
                          uint64_t* codes_out, double* raw_cm_out, int64_t ldr, void* stream);

/* K6  replaces space.get_random_configuration(size=N, use_priors, return_as_array=True) (space.py:1263-1311) followed
 * by the encoding: candidate i of the pool = f(seed, index_base + i, use_priors) (Philox4x32-10; same distributions as
 * the reference's per-parameter randomly_select / randomly_select_uniform, NOT numpy's Mersenne-Twister stream).
 * Any output may be NULL.  prob_out: the prior probability of each candidate (K7) in the same pass.
 * hm_sample_rows re-generates the raw rows [n][D] of the given pool indices (the top-k winners). */
int hm_sample_pool(const hm_space_t* s, uint64_t seed, int64_t index_base, int64_t N, int use_priors,
                   double* raw_cm_out, int64_t ldr, float* X32cm, int64_t ld32, double* X64cm, int64_t ld64,
                   const double* objective_weights_host, int n_objectives, double* prob_out, void* stream);
int hm_sample_rows(const hm_space_t* s, uint64_t seed, int use_priors, const int64_t* index_dev, int64_t n,
                   double* raw_rows_out, void* stream);

/* K7  replaces prior_optimization.compute_probability_from_prior (prior_optimization.py:65-95, the
 * non-estimated branch): prob = prod_params prod_objectives get_x_probability(x) ** w_objective. */
int hm_prior_prob(const hm_space_t* s, const double* raw, int raw_col_major, int64_t ldr, int64_t N,
                  const double* objective_weights_host, int n_objectives, double* prob_out, int32_t* n_invalid_dev,
                  void* stream);
/* piBO weighting (pibo.py:101-113): out = acq * (prob + prior_floor) ** exponent with exponent = prior_beta /
 * iteration_number; collapse_to_min = 1 reproduces the reference's np.min(acq, 0) for TS/UCB (pibo.py:103-104;
 * workspace8 = 8 bytes of device memory). */
int hm_prior_weight(const double* acq, const double* prob, int64_t N, double prior_floor, double exponent,
                    int collapse_to_min, double* out, void* workspace8, void* stream);

/* N4  multi-start hill climb on the device, discrete spaces (get_neighbors, local_search.py:83-161, draws no random
 * numbers there and every configuration has the same number K of neighbours):
 *   hm_ls_neighbors  rows_out[(l*K + q)][D] = neighbour q of live start l (raw values, row-major), enumerated in the
 *                    reference's order; plan_dev = K x {parameter, rank inside that parameter's emitted list} (int32
 *                    pairs, built by the host: local_search.neighbor_plan); current_dev = [S][D] raw configurations;
 *                    live_dev = indices of the starts still climbing
 *   hm_ls_move       per live start: argmin of its K scores in get_min_configurations' order; converged[s] = 1 if the
 *                    best neighbour is the unchanged configuration, else current[s] <- that neighbour; best_q[s] = q */
int hm_ls_neighbors(const hm_space_t* s, const int32_t* plan_dev, int K, const double* current_dev,
                    const int32_t* live_dev, int n_live, double* rows_out, void* stream);
int hm_ls_move(const double* scores, const double* rows, int K, int D, const int32_t* live_dev, int n_live,
               double* current_dev, int32_t* converged_dev, int32_t* best_q_dev, void* stream);

/* BOPrO acquisition (prior_optimization.compute_EI_from_posteriors, prior_optimization.py:161-374):
 *   hm_minmax        min / max of a device vector (NaNs skipped; ignore_inf skips +-inf as :318-327 does);
 *                    out2_dev = {min, max} (+inf / -inf when nothing qualifies); workspace16 = 16 device bytes
 *   hm_bopro_ratio   log posterior(good) - log posterior(bad) per candidate from the prior probability (K7) and the
 *                    per-objective model mean / VARIANCE [M][ldm] (the kernel applies models.py:762-763 and the sqrt)
 *   hm_bopro_finish  optional rescale to [floor, 1] with the running limits, + feasibility indicator, negate,
 *                    +-inf -> +-sys.maxsize; feas_prob = P(feasible) per candidate or NULL */
typedef struct hm_bopro_params {
    int32_t n_objectives;
    int32_t prior_mode;        /* 0 = prior as is, 1 = rescale with [prior_lo, prior_hi] (:184-192), 2 = all ones (:181) */
    int32_t discrete_space;    /* prior_bad /= discrete_divisor (:199-205) */
    int32_t reserved;
    double weight[HM_MAX_OBJECTIVES];
    double threshold[HM_MAX_OBJECTIVES];
    double prior_lo, prior_hi;
    double posterior_floor;
    double discrete_divisor;   /* get_discrete_space_size() - 1 */
    double model_exponent;     /* iteration_number / model_weight */
} hm_bopro_params_t;
int hm_minmax(const double* v, int64_t N, int ignore_inf, double* out2_dev, void* workspace16, void* stream);
int hm_bopro_ratio(const hm_bopro_params_t* p, const double* prior_prob, const double* mean, const double* var,
                   int64_t ldm, int64_t N, double* ratio_out, void* stream);
int hm_bopro_finish(const double* ratio, int64_t N, int normalize_mode, double lo, double hi, double posterior_floor,
                    const double* feas_prob, double* out, double* feas_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer convenience entry points (what a reference-side ctypes/cffi binding calls when its
 * data lives in numpy arrays): pinned staging + H2D/D2H copies happen inside.  Used for `e2e`.
 * ------------------------------------------------------------------------------------------ */
typedef struct hm_host_ctx hm_host_ctx_t;
int hm_host_ctx_create(int64_t max_candidates_per_chunk, int max_features, hm_host_ctx_t** out);
void hm_host_ctx_destroy(hm_host_ctx_t* c);
/* X_host: float32 column-major [n_features][N] (ld = ldx) in host memory (pinned or pageable).
 * Scores all N candidates through M forests + acquisition, chunked and double-buffered across two
 * streams; writes score_out_host[N] if non-NULL and the stable top-k (values + indices) if k > 0. */
int hm_rf_score_host(hm_host_ctx_t* c, const hm_forest_t* const* forests, int M, const float* X_host,
                     int64_t ldx, int64_t N, const hm_acq_params_t* acq, double* score_out_host, int k,
                     double* topk_vals_host, int64_t* topk_idx_host);

/* The same from packed 64-bit codes in host memory (all-discrete spaces; hm_forest_attach_packing done on every forest):
 * 8 bytes per candidate H2D instead of 4 * n_features. */
int hm_rf_score_packed_host(hm_host_ctx_t* c, const hm_forest_t* const* forests, int M, const uint64_t* codes_host,
                            int64_t N, const hm_acq_params_t* acq, double* score_out_host, int k,
                            double* topk_vals_host, int64_t* topk_idx_host);

/* The same from RAW parameter values, row-major fp64 [N][n_params] in host memory -- the form the reference's API is
 * handed (run_acquisition_function's configurations, random_scalarizations.py:75-157): per chunk H2D -> hm_encode ->
 * fused forest + acquisition kernel -> D2H, double-buffered.  *n_invalid_out_host (nullable) receives the number of
 * CTAs that met a discrete value outside its parameter's value list (the reference raises ValueError there). */
int hm_rf_score_raw_host(hm_host_ctx_t* c, const hm_space_t* space, const hm_forest_t* const* forests, int M,
                         const double* raw_host, int64_t N, const hm_acq_params_t* acq, double* score_out_host, int k,
                         double* topk_vals_host, int64_t* topk_idx_host, int32_t* n_invalid_out_host);

/* GP path from host buffers (compute_gp_prediction_mean_and_uncertainty + EI/UCB, models.py:997-1024 +
 * random_scalarizations.py:160-506): X_host = fp64 column-major [n_features][N] (ld = ldx) encoded candidates; per chunk
 * H2D -> hm_gp_mean_var per objective -> hm_acq_score -> D2H, double-buffered; stable top-k if k > 0. */
int hm_gp_score_host(hm_host_ctx_t* c, const hm_gp_t* const* gps, int M, int n_features, const double* X_host,
                     int64_t ldx, int64_t N, const hm_acq_params_t* acq, double* score_out_host, int k,
                     double* topk_vals_host, int64_t* topk_idx_host);

#ifdef __cplusplus
}
#endif
#endif /* HM_B200_H */
