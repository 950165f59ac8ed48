"""Hutter-style forest container + a fit helper for synthetic workloads.

Model fitting is OUT OF SCOPE of the hot path: in a deployment the fitted model is the reference's
`RFModel` (models.py:68-275) produced by `generate_mono_output_regression_models` (models.py:355-499).
bench.py / tests / smoke() run where the reference is not installed, so they need *some* way to obtain a
fitted model of the named shape: `fit_hutter_forest` does what `RFModel.fit_RF` (models.py:184-223) does
-- sklearn fit, per-leaf mean / unbiased variance tables, split thresholds re-drawn uniformly between the
neighbouring training values -- on top of scikit-learn, and returns an object with the same three
attributes the exporter reads (`estimators_`, `tree_means_per_leaf`, `tree_vars_per_leaf`).
"""
import numpy as np


class HutterForest:
    def __init__(self, estimators, means, variances, n_features):
        self.estimators_ = estimators
        self.tree_means_per_leaf = means
        self.tree_vars_per_leaf = variances
        self.n_features_in_ = n_features
        self.n_estimators = len(estimators)


def fit_hutter_forest(X, y, n_estimators=10, max_features=0.5, min_samples_split=5, bootstrap=False, seed=0):
    """X: [n][D_enc] encoded training inputs, y: [n]."""
    from sklearn.ensemble import RandomForestRegressor

    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    rf = RandomForestRegressor(n_estimators=n_estimators, max_features=max_features, bootstrap=bootstrap,
                               min_samples_split=min_samples_split, random_state=seed, n_jobs=None)
    rf.fit(X, y)
    rng = np.random.RandomState(seed + 1)
    X32 = X.astype(np.float32)
    means, variances = [], []
    for est in rf.estimators_:
        tr = est.tree_
        leaf = tr.apply(X32)
        # per-leaf mean and unbiased variance of the training targets (models.py:79-122)
        cnt = np.bincount(leaf, minlength=tr.node_count)
        sm = np.bincount(leaf, weights=y, minlength=tr.node_count)
        lm, lv = {}, {}
        for node in np.nonzero(cnt)[0]:
            ys = y[leaf == node]
            lm[int(node)] = float(sm[node] / cnt[node])
            lv[int(node)] = float(np.var(ys, ddof=1)) if len(ys) > 1 else 0
        means.append(lm)
        variances.append(lv)
        # re-draw every split uniformly between the two training values that straddle it (models.py:202-223)
        path = est.decision_path(X32).tocsc()
        thr = tr.threshold
        for node in range(tr.node_count):
            if tr.children_left[node] == tr.children_right[node]:
                continue
            rows = path.indices[path.indptr[node]:path.indptr[node + 1]]
            vals = X[rows, tr.feature[node]]
            le = vals[vals <= thr[node]]
            gt = vals[vals > thr[node]]
            lo = le.max() if len(le) else -np.inf
            hi = gt.min() if len(gt) else np.inf
            thr[node] = rng.uniform(lo, hi) if np.isfinite(lo) and np.isfinite(hi) else thr[node]
    return HutterForest(list(rf.estimators_), means, variances, X.shape[1])
