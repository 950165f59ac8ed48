/*
 * hm_oracle.c -- CPU restatement (plain C) of the reference's surrogate-prediction / Pareto arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY: it is the checker for the CUDA path (tests/, __graft_entry__.smoke(),
 * bench.py's cpu_baseline / --impl reference legs).  The product never calls it.
 *
 * Pinned against the reference itself: tests/golden/ holds outputs of the unmodified reference
 * (luinardi/hypermapper v2.2.12, run in the build container by tests/golden/make_golden.py) and
 * tests/test_oracle.py checks every function below against them bit for bit.
 *
 * Each function cites the reference lines it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* models.py:174-182 -> sklearn DecisionTreeRegressor.apply -> Tree._apply_dense
 * (sklearn/tree/_tree.pyx): X is cast to float32, the test is `X[i, feature] <= threshold` with a
 * float64 threshold, children == -1 (TREE_LEAF) marks a leaf.  One tree:
 *   X       : float32 row-major [N][D]
 *   out     : int64 [N] leaf node id                                                     */
void oracle_tree_apply(const int32_t* left, const int32_t* right, const int32_t* feature,
                       const double* threshold, const float* X, int64_t N, int D, int64_t* out) {
    for (int64_t i = 0; i < N; ++i) {
        const float* x = X + i * (int64_t)D;
        int32_t node = 0;
        while (left[node] != -1) {
            if ((double)x[feature[node]] <= threshold[node])
                node = left[node];
            else
                node = right[node];
        }
        out[i] = node;
    }
}

/* models.py:181: np.array([tree.apply(bufferx) for tree in self]) -> [T][N] */
void oracle_forest_apply(int T, const int64_t* tree_off, const int32_t* left, const int32_t* right,
                         const int32_t* feature, const double* threshold, const float* X, int64_t N, int D,
                         int64_t* out) {
    for (int t = 0; t < T; ++t) {
        const int64_t b = tree_off[t];
        oracle_tree_apply(left + b, right + b, feature + b, threshold + b, X, N, D, out + (int64_t)t * N);
    }
}

/* models.py:225-240 compute_rf_prediction: tree-outer / sample-inner loops,
 *   sample_means[s] += tree_means_per_leaf[t][leaf] / number_of_trees
 * leaf_mean is indexed [tree_off[t] + leaf id].                                           */
void oracle_rf_prediction(int T, const int64_t* tree_off, const int64_t* leaf /*[T][N]*/, int64_t N,
                          const double* leaf_mean, double* sample_means) {
    for (int64_t s = 0; s < N; ++s) sample_means[s] = 0.0;
    for (int t = 0; t < T; ++t)
        for (int64_t s = 0; s < N; ++s)
            sample_means[s] += leaf_mean[tree_off[t] + leaf[(int64_t)t * N + s]] / (double)T;
}

/* models.py:242-275 compute_rf_prediction_variance (Hutter law of total variance).
 * `x ** 2` on Python/numpy float scalars is libm pow(x, 2.0), restated as such.   (built with -fno-builtin-pow: gcc would fold pow(x,2.0) into x*x, which differs in ~0.09 % of inputs)          */
void oracle_rf_prediction_variance(int T, const int64_t* tree_off, const int64_t* leaf, int64_t N,
                                   const double* leaf_mean, const double* leaf_var, const double* sample_means,
                                   double* sample_vars) {
    for (int64_t s = 0; s < N; ++s) {
        double mov = 0.0, vom = 0.0;
        for (int t = 0; t < T; ++t) {
            const int64_t l = tree_off[t] + leaf[(int64_t)t * N + s];
            mov += leaf_var[l] / (double)T;
            vom += pow(leaf_mean[l], 2.0) / (double)T;
        }
        vom = fabs(vom - pow(sample_means[s], 2.0));
        double v = mov + vom;
        if (v == 0) v = 0.00001;
        sample_vars[s] = v;
    }
}

/* RandomForestRegressor.predict as reached from models.model_prediction (models.py:879-890):
 * sklearn/ensemble/_forest.py accumulates tree.predict(X) in estimator order, then divides by T. */
void oracle_rf_sklearn_predict(int T, const int64_t* tree_off, const int64_t* leaf, int64_t N,
                               const double* node_value, double* out) {
    for (int64_t s = 0; s < N; ++s) out[s] = 0.0;
    for (int t = 0; t < T; ++t)
        for (int64_t s = 0; s < N; ++s) out[s] += node_value[tree_off[t] + leaf[(int64_t)t * N + s]];
    for (int64_t s = 0; s < N; ++s) out[s] /= (double)T;
}

/* Same arithmetic as the three functions above, fused per candidate and parallelised over candidates with
 * OpenMP: the CPU baseline ("port", all host threads).  Per-candidate results are identical to the
 * unfused restatement because each candidate's sums are still taken sequentially in tree order.
 *   X: float32 row-major [N][D];  outputs fp64 [N] (any may be NULL).                     */
void oracle_rf_mean_var_omp(int T, const int64_t* tree_off, const int32_t* left, const int32_t* right,
                            const int32_t* feature, const double* threshold, const double* leaf_mean,
                            const double* leaf_var, const double* node_value, const float* X, int64_t N, int D,
                            double* mean_out, double* var_out, double* pred_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const float* x = X + i * (int64_t)D;
        double mean = 0.0, mov = 0.0, vom = 0.0, pred = 0.0;
        for (int t = 0; t < T; ++t) {
            const int64_t b = tree_off[t];
            int32_t node = 0;
            while (left[b + node] != -1) {
                if ((double)x[feature[b + node]] <= threshold[b + node])
                    node = left[b + node];
                else
                    node = right[b + node];
            }
            const int64_t l = b + node;
            if (leaf_mean) {
                mean += leaf_mean[l] / (double)T;
                mov += leaf_var[l] / (double)T;
                vom += pow(leaf_mean[l], 2.0) / (double)T;
            }
            if (node_value) pred += node_value[l];
        }
        vom = fabs(vom - pow(mean, 2.0));
        double v = mov + vom;
        if (v == 0) v = 0.00001;
        if (mean_out) mean_out[i] = mean;
        if (var_out) var_out[i] = v;
        if (pred_out) pred_out[i] = pred / (double)T;
    }
}

/* compute_pareto.py:89-117 sequential_is_pareto_efficient, both passes, literally:
 *  pass 1: for each still-efficient i (index order): every efficient j with NOT any(c_j <= c_i) dies
 *  pass 2: for each efficient i (index order): dies if some currently efficient j has
 *          any(c_j == c_i) and any(c_j < c_i)
 * costs fp64 row-major [N][M]; mask uint8 [N].                                             */
void oracle_pareto(const double* costs, int64_t N, int M, uint8_t* mask) {
    int64_t* alive = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    int64_t n_alive = N;
    for (int64_t i = 0; i < N; ++i) {
        mask[i] = 1;
        alive[i] = i;
    }
    /* pass 1: `alive` is the compacted list of efficient indices (= costs[is_efficient]) */
    for (int64_t i = 0; i < N; ++i) {
        if (!mask[i]) continue;
        const double* c = costs + i * (int64_t)M;
        int64_t w = 0;
        for (int64_t a = 0; a < n_alive; ++a) {
            const int64_t j = alive[a];
            const double* cj = costs + j * (int64_t)M;
            int any_le = 0;
            for (int d = 0; d < M; ++d) any_le |= (cj[d] <= c[d]);
            if (any_le)
                alive[w++] = j;
            else
                mask[j] = 0;
        }
        n_alive = w;
    }
    /* pass 2 */
    for (int64_t i = 0; i < N; ++i) {
        if (!mask[i]) continue;
        const double* c = costs + i * (int64_t)M;
        int kill = 0;
        for (int64_t a = 0; a < n_alive && !kill; ++a) {
            const int64_t j = alive[a];
            if (!mask[j]) continue;
            const double* cj = costs + j * (int64_t)M;
            int any_eq = 0, any_lt = 0;
            for (int d = 0; d < M; ++d) {
                any_eq |= (cj[d] == c[d]);
                any_lt |= (cj[d] < c[d]);
            }
            kill = any_eq && any_lt;
        }
        if (kill) mask[i] = 0;
    }
    free(alive);
}

/* pass 1 only (compute_pareto.py:100-104) */
void oracle_pareto_pass1(const double* costs, int64_t N, int M, uint8_t* mask) {
    int64_t* alive = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    int64_t n_alive = N;
    for (int64_t i = 0; i < N; ++i) {
        mask[i] = 1;
        alive[i] = i;
    }
    for (int64_t i = 0; i < N; ++i) {
        if (!mask[i]) continue;
        const double* c = costs + i * (int64_t)M;
        int64_t w = 0;
        for (int64_t a = 0; a < n_alive; ++a) {
            const int64_t j = alive[a];
            const double* cj = costs + j * (int64_t)M;
            int any_le = 0;
            for (int d = 0; d < M; ++d) any_le |= (cj[d] <= c[d]);
            if (any_le)
                alive[w++] = j;
            else
                mask[j] = 0;
        }
        n_alive = w;
    }
    free(alive);
}
