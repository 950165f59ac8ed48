"""numpy / ctypes front-end of the CPU oracle (see oracle/__init__.py for status and rules)."""
import ctypes
import os
import subprocess

import numpy as np
from scipy import special as _sp_special
from scipy import linalg as _sp_linalg

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libhm_oracle.so")
_lib = None


def build(force=False):
    """gcc -O2 -fopenmp hm_oracle.c -> oracle/_build/libhm_oracle.so"""
    src = os.path.join(_HERE, "hm_oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(_BUILD, exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-fno-fast-math", "-ffp-contract=off", "-fno-builtin-pow", "-shared", "-fPIC", src, "-o", _SO, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core"""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    lib().oracle_set_num_threads(ctypes.c_int(n))
    return num_threads()


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------------------------
# forest arrays
# ----------------------------------------------------------------------------------------------
class ForestArrays:
    """All trees of one RFModel concatenated, sklearn layout.  leaf_mean / leaf_var are the Hutter
    per-leaf tables of models.py:79-122 scattered to per-node arrays (0 for internal nodes)."""

    def __init__(self, tree_off, left, right, feature, threshold, leaf_mean, leaf_var, value, n_features):
        self.tree_off = np.ascontiguousarray(tree_off, dtype=np.int64)
        self.left = np.ascontiguousarray(left, dtype=np.int32)
        self.right = np.ascontiguousarray(right, dtype=np.int32)
        self.feature = np.ascontiguousarray(feature, dtype=np.int32)
        self.threshold = np.ascontiguousarray(threshold, dtype=np.float64)
        self.leaf_mean = None if leaf_mean is None else np.ascontiguousarray(leaf_mean, dtype=np.float64)
        self.leaf_var = None if leaf_var is None else np.ascontiguousarray(leaf_var, dtype=np.float64)
        self.value = None if value is None else np.ascontiguousarray(value, dtype=np.float64)
        self.n_features = int(n_features)
        self.T = len(self.tree_off) - 1

    def to_npz_dict(self, prefix):
        d = {prefix + k: getattr(self, k) for k in ("tree_off", "left", "right", "feature", "threshold")}
        for k in ("leaf_mean", "leaf_var", "value"):
            if getattr(self, k) is not None:
                d[prefix + k] = getattr(self, k)
        d[prefix + "n_features"] = np.int64(self.n_features)
        return d

    @staticmethod
    def from_npz(z, prefix):
        g = lambda k: z[prefix + k] if (prefix + k) in z else None  # noqa: E731
        return ForestArrays(g("tree_off"), g("left"), g("right"), g("feature"), g("threshold"), g("leaf_mean"),
                            g("leaf_var"), g("value"), int(z[prefix + "n_features"]))


def forest_arrays_from_model(model):
    """model: the reference's RFModel (or anything with .estimators_, .tree_means_per_leaf,
    .tree_vars_per_leaf as filled by fit_RF, models.py:184-223)."""
    offs = [0]
    left, right, feat, thr, val, lm, lv = [], [], [], [], [], [], []
    means = getattr(model, "tree_means_per_leaf", None) or None
    vars_ = getattr(model, "tree_vars_per_leaf", None) or None
    for t, est in enumerate(model.estimators_):
        tr = est.tree_
        n = tr.node_count
        offs.append(offs[-1] + n)
        left.append(np.asarray(tr.children_left, dtype=np.int32))
        right.append(np.asarray(tr.children_right, dtype=np.int32))
        feat.append(np.asarray(tr.feature, dtype=np.int32))
        thr.append(np.asarray(tr.threshold, dtype=np.float64))
        val.append(np.asarray(tr.value, dtype=np.float64).reshape(n, -1)[:, 0])
        if means is not None:
            m = np.zeros(n)
            v = np.zeros(n)
            for leaf, x in means[t].items():
                m[int(leaf)] = x
            for leaf, x in vars_[t].items():
                v[int(leaf)] = x
            lm.append(m)
            lv.append(v)
    cat = np.concatenate
    return ForestArrays(np.array(offs), cat(left), cat(right), cat(feat), cat(thr), cat(lm) if lm else None,
                        cat(lv) if lv else None, cat(val), model.n_features_in_)


# ----------------------------------------------------------------------------------------------
# encoding: models.preprocess_data_array (models.py:941-984)
# ----------------------------------------------------------------------------------------------
def encode(columns, params, normalize_inputs=False, means=None, stds=None):
    """columns: {param name: sequence of raw values (categoricals as integer indices)}
    params: [{"name", "type" in ordinal|real|integer|categorical, "values"}], in input-parameter order.
    Returns (float64 [N][D_enc], list of encoded column names).  Non-categoricals first in input order,
    then each categorical's one-hot block."""
    cols, names = [], []
    for p in params:
        if p["type"] == "categorical":
            continue
        raw = columns[p["name"]]
        if p["type"] == "ordinal":
            vals = list(p["values"])
            size = len(vals)
            cols.append(np.array([vals.index(v) / size for v in raw], dtype=np.float64))
        elif normalize_inputs:
            x = np.array(raw, dtype=np.float64)
            cols.append((x - means[p["name"]]) / stds[p["name"]])
        else:
            cols.append(np.array(raw, dtype=np.float64))
        names.append(p["name"])
    for p in params:
        if p["type"] != "categorical":
            continue
        size = len(p["values"])
        x = np.array(columns[p["name"]]).astype(np.int64)
        for c in range(size):
            cols.append((x == c).astype(np.float64))
            names.append("%s_%d" % (p["name"], c))
    n = len(cols[0]) if cols else 0
    out = np.empty((n, len(cols)), dtype=np.float64)
    for j, c in enumerate(cols):
        out[:, j] = c
    return out, names


# ----------------------------------------------------------------------------------------------
# RF prediction
# ----------------------------------------------------------------------------------------------
def rf_apply(fa, X):
    """[T][N] int64 leaf ids; X is cast to float32 row-major exactly as sklearn's check_array does."""
    X32 = np.ascontiguousarray(X, dtype=np.float32)
    N, D = X32.shape
    out = np.empty((fa.T, N), dtype=np.int64)
    lib().oracle_forest_apply(ctypes.c_int(fa.T), _p(fa.tree_off), _p(fa.left), _p(fa.right), _p(fa.feature),
                              _p(fa.threshold), _p(X32), ctypes.c_int64(N), ctypes.c_int(D), _p(out))
    return out


def rf_prediction(fa, leaf):
    leaf = np.ascontiguousarray(leaf, dtype=np.int64)
    N = leaf.shape[1]
    out = np.empty(N)
    lib().oracle_rf_prediction(ctypes.c_int(fa.T), _p(fa.tree_off), _p(leaf), ctypes.c_int64(N), _p(fa.leaf_mean), _p(out))
    return out


def rf_prediction_variance(fa, leaf, sample_means):
    leaf = np.ascontiguousarray(leaf, dtype=np.int64)
    sm = np.ascontiguousarray(sample_means, dtype=np.float64)
    N = leaf.shape[1]
    out = np.empty(N)
    lib().oracle_rf_prediction_variance(ctypes.c_int(fa.T), _p(fa.tree_off), _p(leaf), ctypes.c_int64(N),
                                        _p(fa.leaf_mean), _p(fa.leaf_var), _p(sm), _p(out))
    return out


def rf_sklearn_predict(fa, leaf):
    leaf = np.ascontiguousarray(leaf, dtype=np.int64)
    N = leaf.shape[1]
    out = np.empty(N)
    lib().oracle_rf_sklearn_predict(ctypes.c_int(fa.T), _p(fa.tree_off), _p(leaf), ctypes.c_int64(N), _p(fa.value), _p(out))
    return out


def rf_mean_var_omp(fa, X, want_pred=False):
    """fused + OpenMP (CPU baseline); identical per-candidate arithmetic."""
    X32 = np.ascontiguousarray(X, dtype=np.float32)
    N, D = X32.shape
    mean = np.empty(N)
    var = np.empty(N)
    pred = np.empty(N) if want_pred else None
    lib().oracle_rf_mean_var_omp(ctypes.c_int(fa.T), _p(fa.tree_off), _p(fa.left), _p(fa.right), _p(fa.feature),
                                 _p(fa.threshold), _p(fa.leaf_mean), _p(fa.leaf_var), _p(fa.value), _p(X32),
                                 ctypes.c_int64(N), ctypes.c_int(D), _p(mean), _p(var), _p(pred))
    return mean, var, pred


def compute_rf_prediction_mean_and_uncertainty(forests, X, var=False):
    """models.py:767-794 on already-encoded X; forests: {objective: ForestArrays}"""
    means, unc = {}, {}
    for obj, fa in forests.items():
        leaf = rf_apply(fa, X)
        means[obj] = rf_prediction(fa, leaf)
        v = rf_prediction_variance(fa, leaf, means[obj])
        unc[obj] = v if var else np.sqrt(v)
    return means, unc


def patch_nan_means(means):
    """models.py:762-763"""
    import sys
    for obj in means:
        means[obj][np.isnan(means[obj])] = sys.maxsize
    return means


# ----------------------------------------------------------------------------------------------
# scalarisation helpers: utility_functions.py:666-756
# ----------------------------------------------------------------------------------------------
def reciprocate_weights(objective_weights):
    new_weights = {}
    total = 0
    for o in objective_weights:
        new_weights[o] = 1 / objective_weights[o]
        total += new_weights[o]
    for o in new_weights:
        new_weights[o] = new_weights[o] / total
    return new_weights


def compute_data_array_scalarization(data_array, objective_weights, objective_limits, scalarization_method):
    import copy
    n = len(data_array[list(data_array.keys())[0]])
    lim = copy.deepcopy(objective_limits)
    norm = {}
    for o in objective_limits:
        lim[o][0] = min(min(data_array[o]), lim[o][0])
        lim[o][1] = max(max(data_array[o]), lim[o][1])
        if objective_limits[o][1] == objective_limits[o][0]:
            norm[o] = [0] * len(data_array[o])
        else:
            norm[o] = [(x - lim[o][0]) / (lim[o][1] - lim[o][0]) for x in data_array[o]]
    if scalarization_method == "linear":
        out = np.zeros(n)
        for r in range(n):
            for o in objective_weights:
                out[r] += objective_weights[o] * norm[o][r]
    elif scalarization_method == "tchebyshev":
        out = np.zeros(n)
        for r in range(n):
            tot = 0
            for o in objective_weights:
                sv = objective_weights[o] * abs(norm[o][r])
                out[r] = max(sv, out[r])
                tot += sv
            out[r] += 0.05 * tot
    elif scalarization_method == "modified_tchebyshev":
        out = np.full(n, float("inf"))
        rw = reciprocate_weights(objective_weights)
        for r in range(n):
            for o in objective_weights:
                sv = rw[o] * abs(norm[o][r])
                out[r] = min(sv, out[r])
            out[r] = -out[r]
    else:
        raise SystemExit
    return out, lim


# ----------------------------------------------------------------------------------------------
# acquisition functions on predicted mean / variance: random_scalarizations.py:160-506
# ----------------------------------------------------------------------------------------------
def _norm_cdf(v):
    return _sp_special.ndtr(v)   # scipy.stats.norm.cdf == ndtr for loc=0, scale=1


def _norm_pdf(v):
    return np.exp(-v ** 2 / 2.0) / np.sqrt(2 * np.pi)   # scipy.stats.norm._pdf


def ei_f_min(data_array, objective, lim):
    """random_scalarizations.py:419-431"""
    if lim[objective][1] - lim[objective][0] == 0:
        diff = 1
    else:
        diff = lim[objective][1] - lim[objective][0]
    return 1 - (min(data_array[objective]) - lim[objective][0]) / diff


def EI(means, variances, data_array, objective_weights, scalarization_method, objective_limits, feasibility=None):
    objectives = list(means.keys())
    n = len(means[objectives[0]])
    feas = np.ones(n) if feasibility is None else np.asarray(feasibility)
    _, lim = compute_data_array_scalarization(data_array, objective_weights, objective_limits, scalarization_method)

    def term(o):
        f_min = ei_f_min(data_array, o, lim)
        x_std = np.sqrt(variances[o])
        x_mean = 1 - means[o]
        with np.errstate(all="ignore"):
            v = (x_mean - f_min) / x_std
            return (x_mean - f_min) * _norm_cdf(v) + x_std * _norm_pdf(v)

    if scalarization_method == "linear":
        s = np.zeros(n)
        for o in objectives:
            s += term(o) * objective_weights[o]
        return -1 * s * feas
    if scalarization_method == "tchebyshev":
        sp = np.zeros(n)
        tot = np.zeros(n)
        for o in objectives:
            sv = objective_weights[o] * term(o)
            sp = np.maximum(sv, sp)
            tot += sv
        return -1 * (sp + tot * 0.05) * feas
    if scalarization_method == "modified_tchebyshev":
        sp = np.full(n, float("inf"))
        rw = reciprocate_weights(objective_weights)
        for o in objectives:
            sv = rw[o] * term(o)
            sp = np.minimum(sv, sp)
        return sp * feas
    raise SystemExit


def ucb(means, variances, objective_weights, scalarization_method, iteration_number, feasibility=None):
    objectives = list(means.keys())
    n = len(means[objectives[0]])
    feas = np.ones(n) if feasibility is None else np.asarray(feasibility)
    beta = np.sqrt(0.125 * np.log(2 * iteration_number + 1))
    if scalarization_method == "linear":
        s = np.zeros(n)
        bf = 0
        for o in objectives:
            s += objective_weights[o] * means[o]
            bf += objective_weights[o] * variances[o]
        s -= beta * np.sqrt(bf)
        return s * feas
    if scalarization_method == "tchebyshev":
        sp = np.zeros(n)
        tot = np.zeros(n)
        for o in objectives:
            sv = objective_weights[o] * (means[o] - beta * np.sqrt(variances[o]))
            tot += sv
            sp = np.maximum(sv, sp)
        sp += 0.05 * tot
        return sp * feas
    if scalarization_method == "modified_tchebyshev":
        sp = np.full(n, float("inf"))
        rw = reciprocate_weights(objective_weights)
        for o in objectives:
            sv = rw[o] * (means[o] - beta * np.sqrt(variances[o]))
            sp = np.minimum(sv, sp)
        sp = sp * feas
        return -sp
    raise SystemExit


def thompson_sampling(predictions, objective_weights, scalarization_method, feasibility=None):
    objectives = list(predictions.keys())
    n = len(predictions[objectives[0]])
    feas = np.ones(n) if feasibility is None else np.asarray(feasibility)
    if scalarization_method == "linear":
        s = np.zeros(n)
        for o in objectives:
            s += objective_weights[o] * predictions[o]
        return s * feas
    if scalarization_method == "tchebyshev":
        sp = np.zeros(n)
        tot = np.zeros(n)
        for o in objectives:
            sv = objective_weights[o] * np.absolute(predictions[o])
            tot += sv
            sp = np.maximum(sv, sp)
        sp += 0.05 * tot
        return sp * feas
    if scalarization_method == "modified_tchebyshev":
        sp = np.full(n, float("inf"))
        rw = reciprocate_weights(objective_weights)
        for o in objectives:
            sv = rw[o] * np.absolute(predictions[o])
            sp = np.minimum(sv, sp)
        sp = sp * feas
        return -sp
    raise SystemExit


def rf_acquisition(acquisition_function, forests, X, objective_weights, scalarization_method, objective_limits,
                   iteration_number, data_array, feasibility=None):
    """run_acquisition_function (random_scalarizations.py:75-157) for model_type == random_forest on
    already-encoded candidates X."""
    if acquisition_function == "TS":
        preds = {o: rf_sklearn_predict(fa, rf_apply(fa, X)) for o, fa in forests.items()}
        return thompson_sampling(preds, objective_weights, scalarization_method, feasibility)
    means, variances = compute_rf_prediction_mean_and_uncertainty(forests, X, var=True)
    patch_nan_means(means)
    if acquisition_function == "UCB":
        return ucb(means, variances, objective_weights, scalarization_method, iteration_number, feasibility)
    if acquisition_function == "EI":
        return EI(means, variances, data_array, objective_weights, scalarization_method, objective_limits, feasibility)
    raise SystemExit


# ----------------------------------------------------------------------------------------------
# get_min_configurations (utility_functions.py:295-316) in index form
# ----------------------------------------------------------------------------------------------
def get_min_indices(values, k):
    """k rounds of np.argmin + delete, returning ORIGINAL indices (ascending value, ties -> lowest index)."""
    vals = list(values)
    idx = list(range(len(vals)))
    k = min(k, len(vals))
    out = []
    for _ in range(k):
        j = int(np.argmin(vals))
        out.append(idx[j])
        del vals[j]
        del idx[j]
    return np.array(out, dtype=np.int64)


def get_min_indices_fast(values, k):
    """same result for large N (stable argsort on NaN-first, -0==+0 keys)."""
    v = np.asarray(values, dtype=np.float64)
    k = min(k, len(v))
    nan = np.isnan(v)
    key = np.where(nan, -np.inf, v + 0.0)
    # NaN strictly before -inf: lexsort on (index, key, not-nan)
    order = np.lexsort((np.arange(len(v)), key, ~nan))
    return order[:k].astype(np.int64)


# ----------------------------------------------------------------------------------------------
# Pareto: compute_pareto.py:89-117
# ----------------------------------------------------------------------------------------------
def pareto(costs):
    c = np.ascontiguousarray(costs, dtype=np.float64)
    N, M = c.shape
    mask = np.empty(N, dtype=np.uint8)
    lib().oracle_pareto(_p(c), ctypes.c_int64(N), ctypes.c_int(M), _p(mask))
    return mask.astype(bool)


def pareto_pass1(costs):
    c = np.ascontiguousarray(costs, dtype=np.float64)
    N, M = c.shape
    mask = np.empty(N, dtype=np.uint8)
    lib().oracle_pareto_pass1(_p(c), ctypes.c_int64(N), ctypes.c_int(M), _p(mask))
    return mask.astype(bool)


def pareto_numpy(costs):
    """literal numpy restatement (small inputs)"""
    costs = np.asarray(costs, dtype=np.float64)
    eff = np.ones(costs.shape[0], dtype=bool)
    for i, c in enumerate(costs):
        if eff[i]:
            eff[eff] = np.any(costs[eff] <= c, axis=1)
    for i, c in enumerate(costs):
        if eff[i]:
            if np.any(np.logical_and(np.any(costs[eff] == c, axis=1), np.any(costs[eff] < c, axis=1))):
                eff[i] = False
    return eff


# ----------------------------------------------------------------------------------------------
# GP (PARITY UNPINNED -- restates GPy 1.9.9 exact inference; see oracle/__init__.py)
# ----------------------------------------------------------------------------------------------
def _unscaled_dist(A, B):
    """GPy kern/src/stationary.py Stationary._unscaled_dist: GEMM form, clipped at 0, sqrt."""
    a2 = np.sum(np.square(A), 1)
    b2 = np.sum(np.square(B), 1)
    r2 = -2.0 * A.dot(B.T) + (a2[:, None] + b2[None, :])
    if A is B:
        np.fill_diagonal(r2, 0.0)
    r2 = np.clip(r2, 0, np.inf)
    return np.sqrt(r2)


def matern52(A, B, lengthscale, variance):
    """GPy.kern.Matern52(ARD=True).K : sf2 * (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r), r on inputs / lengthscale"""
    As = A / lengthscale
    Bs = As if B is A else B / lengthscale
    r = _unscaled_dist(As, Bs)
    return variance * (1 + np.sqrt(5.0) * r + 5.0 / 3 * r ** 2) * np.exp(-np.sqrt(5.0) * r)


def rbf(A, B, lengthscale, variance):
    As = A / lengthscale
    Bs = As if B is A else B / lengthscale
    r = _unscaled_dist(As, Bs)
    return variance * np.exp(-0.5 * r ** 2)


class GPState:
    pass


def gp_fit(X, y, lengthscale, variance, noise, kernel="matern52"):
    """GPRegression(X, y, Matern52(ARD), normalizer=True) with FIXED hyper-parameters (fitting them is out of
    scope): Standardize normaliser, K_y = K + (noise + 1e-8) I (GPy exact_gaussian_inference adds 1e-8 jitter
    via the likelihood variance + pdinv), L = chol(K_y), alpha = K_y^-1 y_norm (dpotrs)."""
    st = GPState()
    st.X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    st.y_mean = float(y.mean())
    st.y_std = float(y.std())
    yn = (y - st.y_mean) / st.y_std
    st.lengthscale = np.asarray(lengthscale, dtype=np.float64)
    st.variance = float(variance)
    st.noise = float(noise)
    st.kernel = kernel
    kfun = matern52 if kernel == "matern52" else rbf
    K = kfun(st.X, st.X, st.lengthscale, st.variance)
    Ky = K + (st.noise + 1e-8) * np.eye(len(y))
    st.L = np.linalg.cholesky(Ky)
    st.alpha = _sp_linalg.cho_solve((st.L, True), yn)
    return st


def gp_predict(st, Xstar):
    """GPy Posterior._raw_predict + Gaussian likelihood + normaliser inverse, then models.py:1013-1018."""
    Xstar = np.asarray(Xstar, dtype=np.float64)
    kfun = matern52 if st.kernel == "matern52" else rbf
    Kx = kfun(st.X, Xstar, st.lengthscale, st.variance)          # n x N
    mu = Kx.T.dot(st.alpha)
    tmp = _sp_linalg.solve_triangular(st.L, Kx, lower=True)      # dtrtrs
    var = st.variance - np.sum(np.square(tmp), 0)
    var = np.clip(var, 1e-15, np.inf)
    var = var + st.noise
    mean = mu * st.y_std + st.y_mean
    var = var * st.y_std ** 2
    var[var < 10 ** -11] = 10 ** -11
    return mean, var
