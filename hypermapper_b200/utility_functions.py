"""Host-side helpers of the hot path, mirroring the names in the reference's utility_functions.py.

Only what the path needs: dict-of-lists plumbing used by local_search, the observed-data scalarisation whose
limits EI reads (utility_functions.py:666-756), and the k-smallest selection (utility_functions.py:295-366)
which runs on the GPU (hm_topk) for device-resident score vectors.
"""
import copy
from collections import defaultdict

import numpy as np


# ---- dict-of-lists plumbing (utility_functions.py:228-274, 544-617) -------------------------------------
def concatenate_data_dictionaries(D1, D2, selection_keys_list=[]):
    keys = set(selection_keys_list) if len(selection_keys_list) else set(list(D1.keys()) + list(D2.keys()))
    out = {}
    for k in keys:
        if len(D1) > 0 and len(D2) > 0:
            out[k] = D1[k][:] + D2[k][:]
        elif len(D1) == 0:
            out[k] = D2[k][:]
        else:
            out[k] = D1[k][:]
    return out


def concatenate_list_of_dictionaries(dict_list, selection_keys_list=None):
    out = defaultdict(list)
    if selection_keys_list is None:
        selection_keys_list = list(dict_list[0].keys())
    for d in dict_list:
        for k in selection_keys_list:
            out[k].append(d[k])
    return out


def get_single_configuration(configurations, idx):
    return {k: configurations[k][idx] for k in configurations}


def are_configurations_equal(configuration1, configuration2, keys):
    return all(configuration1[k] == configuration2[k] for k in keys)


def data_tuples_to_dict_list(data_tuple, keys):
    return [{k: entry[j] for j, k in enumerate(keys)} for entry in data_tuple]


def data_dictionary_to_tuple(data_dict, keys):
    return [tuple(data_dict[k][i] for k in keys) for i in range(len(data_dict[keys[0]]))]


def dict_of_lists_to_list_of_dicts(data_dict, keys=None):
    if keys is None:
        keys = list(data_dict.keys())
    return [{k: data_dict[k][i] for k in keys} for i in range(len(data_dict[keys[0]]))]


def array_to_list_of_dicts(data_array, keys):
    return [{k: data_array[r, j] for j, k in enumerate(keys)} for r in range(len(data_array[:, 0]))]


def dict_of_lists_to_numpy(input_dict, return_col_of_key):
    names = list(input_dict.keys())
    arr = np.zeros((len(input_dict[names[0]]), len(input_dict)), dtype=np.dtype(object))
    col_of_key = {}
    for j, n in enumerate(names):
        col_of_key[n] = j
        arr[:, j] = input_dict[n]
    return (arr, col_of_key) if return_col_of_key else arr


# ---- scalarisation of the observed data (utility_functions.py:666-756) -----------------------------------
def reciprocate_weights(objective_weights):
    inv = {o: 1 / objective_weights[o] for o in objective_weights}
    total = 0
    for o in inv:
        total += inv[o]
    return {o: inv[o] / total for o in inv}


def compute_data_array_scalarization(data_array, objective_weights, objective_limits, scalarization_method):
    """Returns (scalarised observations, objective limits widened by the observed min/max).
    The limits are what EI turns into f_min (random_scalarizations.py:411-431), so they are reproduced exactly."""
    n = len(data_array[list(data_array.keys())[0]])
    lim = copy.deepcopy(objective_limits)
    normed = {}
    for o in objective_limits:
        col = data_array[o]
        lim[o][0] = min(min(col), lim[o][0])
        lim[o][1] = max(max(col), lim[o][1])
        if objective_limits[o][1] == objective_limits[o][0]:   # sic: the *incoming* limits decide (utility_functions.py:709)
            normed[o] = np.zeros(len(col))
        else:
            normed[o] = (np.asarray(col, dtype=np.float64) - lim[o][0]) / (lim[o][1] - lim[o][0])
    if scalarization_method == "linear":
        out = np.zeros(n)
        for o in objective_weights:
            out = out + objective_weights[o] * normed[o]
    elif scalarization_method == "tchebyshev":
        out = np.zeros(n)
        total = np.zeros(n)
        for o in objective_weights:
            sv = objective_weights[o] * np.abs(normed[o])
            out = np.maximum(sv, out)
            total = total + sv
        out = out + 0.05 * total
    elif scalarization_method == "modified_tchebyshev":
        out = np.full(n, float("inf"))
        rw = reciprocate_weights(objective_weights)
        for o in objective_weights:
            out = np.minimum(rw[o] * np.abs(normed[o]), out)
        out = -out
    else:
        print("Error: unrecognized scalarization method:", scalarization_method)
        raise SystemExit
    return out, lim


# ---- k smallest (utility_functions.py:295-366) -----------------------------------------------------------
def min_indices(values, k):
    """Indices of the k smallest entries in the order get_min_configurations returns them (ascending value, ties ->
    lowest index, NaN first).  CUDA tensors go through hm_topk; host sequences of any size use the same ordering
    via a stable host sort (used for the tiny neighbour batches of the hill climb)."""
    try:
        import torch
        if isinstance(values, torch.Tensor) and values.is_cuda:
            from .forest import topk
            k = min(int(k), values.numel())
            if k == 0:
                return np.zeros(0, dtype=np.int64)
            v = values.contiguous().view(-1).to(torch.float64)
            if k <= 1024:
                _, idx = topk(v, k)
                return idx.cpu().numpy()
            # beyond the k <= 1024 limit of hm_topk (a pool whose best 256+ entries were all evaluated already, or
            # local_search_starting_points > 512): full stable device sort on the same key
            nan = torch.isnan(v)
            order = torch.sort(torch.where(nan, torch.full_like(v, float("-inf")), v + 0.0), stable=True).indices
            isn = nan[order]
            return torch.cat([order[isn], order[~isn]])[:k].cpu().numpy()
    except ImportError:
        pass
    v = np.asarray(values, dtype=np.float64)
    k = min(int(k), len(v))
    nan = np.isnan(v)
    key = np.where(nan, -np.inf, v + 0.0)
    return np.lexsort((np.arange(len(v)), key, ~nan))[:k].astype(np.int64)


def get_min_configurations(configurations, number_of_configurations, comparison_key):
    """dict of lists -> dict of lists holding the best `number_of_configurations` rows."""
    size = len(configurations[list(configurations.keys())[0]])
    k = min(number_of_configurations, size)
    idx = min_indices(configurations[comparison_key], k)
    best = defaultdict(list)
    for key in configurations:
        col = configurations[key]
        for i in idx:
            best[key].append(col[int(i)])
    return best


def get_min_feasible_configurations(configurations, number_of_configurations, comparison_key, feasible_parameter):
    """utility_functions.py:319-366: feasible rows first (all of them, in order, if there are not enough), topped up
    with the best infeasible ones; the best feasible ones if there are too many."""
    size = len(configurations[list(configurations.keys())[0]])
    k = min(number_of_configurations, size)
    feas_idx = [i for i in range(size) if configurations[feasible_parameter][i]]
    infe_idx = [i for i in range(size) if not configurations[feasible_parameter][i]]

    def take(rows):
        return {key: [configurations[key][i] for i in rows] for key in configurations}

    if len(feas_idx) < k:
        infeasible = take(infe_idx)
        best_inf = get_min_configurations(infeasible, k - len(feas_idx), comparison_key) if infe_idx else {}
        return concatenate_data_dictionaries(take(feas_idx) if feas_idx else {}, best_inf)
    if len(feas_idx) > k:
        return get_min_configurations(take(feas_idx), k, comparison_key)
    return take(feas_idx)
