"""Bit-identical selection: host re-score of near-tied finalists with the reference's own arithmetic.

The device scores agree with the reference to ~1e-13 relative, not bit for bit: CUDA's erf / erfc / exp differ from
scipy's Cephes `ndtr` in the last ulps, and the Hutter variance uses x*x where the reference calls libm pow
(models.py:262-264).  `get_min_configurations` (utility_functions.py:295-316), the hill-climb move
(local_search.py:367-369) and the final pick (local_search.py:789-807) are argmins over those values, so two
candidates whose values differ by a few ulps could be ordered differently.  Candidates that sit in the same leaves of
every tree have bit-identical values on both sides (ties are the norm on discrete spaces; both sides break them by
lowest index), so only DISTINCT values closer than NEAR_TIE_REL need a second look:

  1. the GPU ranks all N candidates (hm_rf_eval + hm_topk);
  2. the few distinct device values inside the tolerance window around the k-th best are re-scored HERE, from the
     leaf ids the GPU reports (integers, bit-exact), with the reference's formula: dict-of-leaf tables, scalar `**`,
     scipy.stats.norm.cdf / pdf, numpy sqrt (random_scalarizations.py:160-506, models.py:225-275);
  3. the selection is redone on (exact value, index).

This is O(dozens) of rows per call; every one of the N scores still comes from the GPU.
"""
import copy
import sys

import numpy as np
from scipy import stats

from .utility_functions import compute_data_array_scalarization, reciprocate_weights

# two device values closer than this (relative) are considered "possibly ordered differently by the reference"
NEAR_TIE_REL = 1e-9
# more distinct values than this inside one tolerance window: keep the device order (never seen; logged in stats)
MAX_RESCORE = 512

stats_counters = {"windows": 0, "rescored_values": 0, "order_changes": 0, "overflow": 0}


def rf_mean_var_from_leaves(model, leaves):
    """models.py:225-275 on a host leaf matrix [T][n]: (mean, variance), the reference's arithmetic and order."""
    from .models import rf_compute_rf_prediction, rf_compute_rf_prediction_variance
    mean = rf_compute_rf_prediction(model, leaves)
    return mean, rf_compute_rf_prediction_variance(model, leaves, mean)


def rf_predict_from_leaves(model, leaves):
    """RandomForestRegressor.predict (sklearn/ensemble/_forest.py: sum of tree_.value[leaf] in tree order, then / T)."""
    leaves = np.asarray(leaves)
    out = np.zeros(leaves.shape[1], dtype=np.float64)
    for t, est in enumerate(model.estimators_):
        out += np.asarray(est.tree_.value, dtype=np.float64).reshape(est.tree_.node_count, -1)[:, 0][leaves[t]]
    out /= len(model.estimators_)
    return out


def reference_acquisition(acquisition_function, scalarization_method, objectives, means, variances, objective_weights,
                          objective_limits, iteration_number, data_array, feasibility=None):
    """random_scalarizations.py:160-506 restated operation for operation on host arrays (n = a handful of rows).
    means / variances: {objective: float64 [n]} (for TS `means` holds the plain forest prediction)."""
    n = len(means[objectives[0]])
    feas = [1] * n if feasibility is None else feasibility
    if acquisition_function == "TS":
        if scalarization_method == "linear":
            out = np.zeros(n)
            for o in objectives:
                out += objective_weights[o] * means[o]
            return out * feas
        if scalarization_method == "tchebyshev":
            out = np.zeros(n)
            total = np.zeros(n)
            for o in objectives:
                sv = objective_weights[o] * np.absolute(means[o])
                total += sv
                out = np.maximum(sv, out)
            out += 0.05 * total
            return out * feas
        if scalarization_method == "modified_tchebyshev":
            out = np.full((n), float("inf"))
            rw = reciprocate_weights(objective_weights)
            for o in objectives:
                out = np.minimum(rw[o] * np.absolute(means[o]), out)
            out = out * feas
            return -out
    elif acquisition_function == "UCB":
        beta = np.sqrt(0.125 * np.log(2 * iteration_number + 1))
        if scalarization_method == "linear":
            out = np.zeros(n)
            beta_factor = 0
            for o in objectives:
                out += objective_weights[o] * means[o]
                beta_factor += objective_weights[o] * variances[o]
            out -= beta * np.sqrt(beta_factor)
            return out * feas
        if scalarization_method == "tchebyshev":
            out = np.zeros(n)
            total = np.zeros(n)
            for o in objectives:
                sv = objective_weights[o] * (means[o] - beta * np.sqrt(variances[o]))
                total += sv
                out = np.maximum(sv, out)
            out += 0.05 * total
            return out * feas
        if scalarization_method == "modified_tchebyshev":
            out = np.full((n), float("inf"))
            rw = reciprocate_weights(objective_weights)
            for o in objectives:
                out = np.minimum(rw[o] * (means[o] - beta * np.sqrt(variances[o])), out)
            out = out * feas
            return -out
    elif acquisition_function == "EI":
        _, lim = compute_data_array_scalarization(data_array, objective_weights, copy.deepcopy(objective_limits),
                                                  scalarization_method)

        def objective_ei(o):
            diff = lim[o][1] - lim[o][0]
            if diff == 0:
                diff = 1
            f_min = 1 - (min(data_array[o]) - lim[o][0]) / (diff)
            x_std = np.sqrt(variances[o])
            x_mean = 1 - means[o]
            with np.errstate(all="ignore"):
                v = (x_mean - f_min) / x_std
                return (x_mean - f_min) * stats.norm.cdf(v) + x_std * stats.norm.pdf(v)

        if scalarization_method == "linear":
            out = np.zeros(n)
            for o in objectives:
                out += objective_ei(o) * objective_weights[o]
            return -1 * out * feas
        if scalarization_method == "tchebyshev":
            out = np.zeros(n)
            total = np.zeros(n)
            for o in objectives:
                sv = objective_weights[o] * objective_ei(o)
                out = np.maximum(sv, out)
                total += sv
            return -1 * (out + total * 0.05) * feas
        if scalarization_method == "modified_tchebyshev":
            out = np.full((n), float("inf"))
            rw = reciprocate_weights(objective_weights)
            for o in objectives:
                out = np.minimum(rw[o] * objective_ei(o), out)
            return out * feas
    else:
        print("Unrecognized acquisition function:", acquisition_function)
        raise SystemExit
    print("Error: unrecognized scalarization method:", scalarization_method)
    raise SystemExit


def reference_rf_scores_from_leaves(acquisition_function, scalarization_method, regression_models, leaves,
                                    objective_weights, objective_limits, iteration_number, data_array,
                                    feasibility=None):
    """Exact reference value of run_acquisition_function (random forest) for the candidates whose leaf matrices
    (one [T][n] per objective, in regression_models order) are given."""
    objectives = list(regression_models.keys())
    means, variances = {}, {}
    for o, lv in zip(objectives, leaves):
        if acquisition_function == "TS":
            means[o] = rf_predict_from_leaves(regression_models[o], lv)
            variances[o] = None
        else:
            means[o], variances[o] = rf_mean_var_from_leaves(regression_models[o], lv)
            means[o][np.isnan(means[o])] = sys.maxsize           # models.py:762-763
    return np.asarray(reference_acquisition(acquisition_function, scalarization_method, objectives, means, variances,
                                            objective_weights, objective_limits, iteration_number, data_array,
                                            feasibility), dtype=np.float64)


# ---- selection ------------------------------------------------------------------------------------------------------
def _key(v):
    """get_min_configurations order on host floats: NaN first, -0.0 == +0.0"""
    return (0, 0.0) if v != v else (1, v + 0.0)


def near_tie_clusters(sorted_values, rel=NEAR_TIE_REL):
    """sorted_values: ascending distinct finite values.  Returns the list of index ranges (lo, hi) (hi exclusive, hi-lo>=2)
    in which consecutive values are closer than rel * max(|a|, |b|)."""
    out = []
    lo = 0
    v = sorted_values
    for i in range(1, len(v) + 1):
        joined = i < len(v) and (v[i] - v[i - 1]) <= rel * max(abs(v[i]), abs(v[i - 1]))
        if not joined:
            if i - lo >= 2:
                out.append((lo, i))
            lo = i
    return out


def exact_min_indices(values, k, exact_of_rows):
    """Indices (np.int64 [k]) of the k smallest entries of the CUDA fp64 vector `values` in the order the REFERENCE's
    get_min_configurations would return them on ITS values: ascending value, ties -> lowest index, NaN first.
    exact_of_rows(idx: np.int64 array) -> np.float64 array of exact reference values of those candidates."""
    import torch
    from .forest import topk
    from .utility_functions import min_indices
    N = values.numel()
    k = min(int(k), N)
    if k == 0:
        return np.zeros(0, dtype=np.int64)
    base = min_indices(values, k)
    if exact_of_rows is None:
        return base
    tv = values[torch.as_tensor(base, device=values.device)]
    v_k = float(tv[-1])
    if v_k != v_k or v_k == float("inf") or v_k == float("-inf"):
        return base
    thr = v_k + NEAR_TIE_REL * abs(v_k)
    window = torch.nonzero((values <= thr) | torch.isnan(values)).view(-1)          # ascending index
    cv = values[window] + 0.0                                                       # -0.0 -> +0.0
    uniq, inv = torch.unique(cv, return_inverse=True)                               # NaNs sort last, each its own entry
    uh = uniq.cpu().numpy()
    finite = ~np.isnan(uh)
    clusters = near_tie_clusters(uh[finite])
    if not clusters:
        return base
    stats_counters["windows"] += 1
    members = sum(hi - lo for lo, hi in clusters)
    if members > MAX_RESCORE:
        stats_counters["overflow"] += 1
        return base
    # representative (lowest index) of every distinct device value that sits in a cluster
    first = torch.full((uniq.numel(),), N, dtype=torch.int64, device=values.device)
    first.scatter_reduce_(0, inv, window, reduce="amin")
    fh = first.cpu().numpy()
    sel = np.concatenate([np.arange(lo, hi) for lo, hi in clusters])
    exact = np.asarray(exact_of_rows(fh[sel]), dtype=np.float64)
    stats_counters["rescored_values"] += len(sel)
    fixed = uh.copy()
    fixed[sel] = exact
    ev = torch.from_numpy(fixed).to(values.device)[inv].contiguous()
    _, pos = topk(ev, k)
    out = window[pos].cpu().numpy().astype(np.int64)
    if not np.array_equal(out, base):
        stats_counters["order_changes"] += 1
    return out


def exact_argmin_host(values, exact_of_positions):
    """First index of the minimum of a host list in the reference's order, re-scoring the distinct values that are
    within NEAR_TIE_REL of the minimum.  exact_of_positions(list of positions) -> exact values."""
    n = len(values)
    best = min(range(n), key=lambda i: _key(values[i]) + (i,))
    vb = values[best]
    if exact_of_positions is None or vb != vb or vb in (float("inf"), float("-inf")):
        return best
    thr = vb + NEAR_TIE_REL * abs(vb)
    near = [i for i in range(n) if values[i] <= thr]
    distinct = {}
    for i in near:
        distinct.setdefault(values[i] + 0.0, i)
    if len(distinct) < 2 or len(distinct) > MAX_RESCORE:
        return best
    stats_counters["windows"] += 1
    reps = list(distinct.values())
    exact = dict(zip(distinct.keys(), np.asarray(exact_of_positions(reps), dtype=np.float64).tolist()))
    stats_counters["rescored_values"] += len(reps)
    pick = min(near, key=lambda i: _key(exact[values[i] + 0.0]) + (i,))
    if pick != best:
        stats_counters["order_changes"] += 1
    return pick
