"""Acquisition optimiser of the hot path: same entry points as the reference's local_search.py
(get_neighbors :83-161, local_search :389-833), with the two hot loops on the GPU:

  A. the two big random pools are encoded as arrays, scored in one fused launch each and reduced to the
     2*starting_points best seeds by the stable device top-k -- no per-candidate Python dict is built
     (the reference explodes the (N, D) sample array into N dicts, local_search.py:446-451);
  B. the hill climb scores every neighbour batch through the same fused kernel with the model resident.

`optimization_function` keeps the reference's plug-in contract
    f(configurations=[dict, ...], **optimization_function_parameters) -> (values, feasibility_indicators)
(local_search.py:468-473, 346-348); any callable works.  When it is this package's run_acquisition_function the
pools bypass the list-of-dicts form.  The multi-process branches of the reference (number_of_cpus != 1) are not
reproduced: a CUDA context must not be forked, and the GPU replaces the worker pool.

Known non-determinism of the reference: the column order used by the duplicate filter comes from iterating a Python
`set` of key strings (utility_functions.py:237-239 via local_search.py:666-675), which changes with the interpreter's
string-hash seed.  Input-parameter order (scalarization key last) is used here; `set_dedup_column_order` pins any other
order, e.g. the one a recorded reference run used (tests/golden/local_search_traj.npz).

Selection identity: the seeds, every hill-climb move and the final pick are argmins over acquisition values; where the
scorer provides `exact_scores` (run_acquisition_function on random forests) the distinct device values that are closer
than 1e-9 relative are re-scored on the host with the reference's own arithmetic before they are ordered (exact.py).
"""
import copy
import datetime
import os
import sys

import numpy as np
from scipy import stats

from .exact import NEAR_TIE_REL, exact_argmin_host, exact_min_indices
from .utility_functions import (
    are_configurations_equal,
    concatenate_data_dictionaries,
    concatenate_list_of_dictionaries,
    data_tuples_to_dict_list,
    dict_of_lists_to_list_of_dicts,
    get_min_configurations,
    get_min_feasible_configurations,
    min_indices,
)

# pools larger than this are not materialised as Python lists in the returned data_array
MATERIALISE_LIMIT = 262144

# Device pools: when the acquisition plug-in is one of this package's device scorers, the two
# random pools can be generated AND encoded on the GPU (hm_sample_pool: same per-parameter distributions as
# space.get_random_configuration, Philox stream instead of numpy's) and never exist on the host; the few rows the
# optimiser needs as configurations are re-generated from (seed, index).
#   "auto" (default): host pools -- drawn by param_space.get_random_configuration from numpy's global stream exactly as
#                     the reference does, so seeded runs pick the reference's configurations -- up to
#                     DEVICE_POOL_AUTO_POINTS candidates per pool; device pools above (the reference's list-of-dicts
#                     path cannot run there, and building the numpy object array alone takes seconds: 2.4 s per
#                     iteration at 2 x 10^6 candidates against 0.03 s)
#   True / False    : always / never       (also HM_B200_DEVICE_POOLS=1 / 0)
DEVICE_POOL_AUTO_POINTS = 1 << 20
_env = os.environ.get("HM_B200_DEVICE_POOLS", "auto")
_device_pools = "auto" if _env == "auto" else (_env not in ("", "0"))


def set_device_pools(mode):
    """True / False / "auto" """
    global _device_pools
    _device_pools = mode if mode == "auto" else bool(mode)


def _use_device_pools(n_points):
    if _device_pools == "auto":
        return n_points >= DEVICE_POOL_AUTO_POINTS
    return bool(_device_pools)


_dedup_column_order = None


def set_dedup_column_order(keys):
    """None (input-parameter order + scalarization key) or the explicit list of column keys the duplicate filter of
    local_search.py:673-699 should see (what iterating the reference's key `set` gave in a recorded run)."""
    global _dedup_column_order
    _dedup_column_order = None if keys is None else list(keys)


def _pool_rows_matrix(pool, inputs, rows):
    """raw parameter values [n][D] fp64 of the given pool rows (categoricals as indices)"""
    rows = [int(r) for r in rows]
    if hasattr(pool, "candidates"):
        return pool.ds.sample_rows(pool.seed, pool.use_priors, np.asarray(rows, dtype=np.int64))
    if isinstance(pool, list):
        return np.array([[float(pool[r][p]) for p in inputs] for r in rows], dtype=np.float64).reshape(len(rows), len(inputs))
    return np.asarray(pool[rows], dtype=np.float64).reshape(len(rows), len(inputs))


def _log(msg, verbose=False):
    w = getattr(sys.stdout, "write_to_logfile", None)
    if w is not None:
        try:
            w(msg, msg_is_verbose=verbose) if verbose else w(msg)
        except TypeError:
            w(msg)


def _truncnorm_draws(means, n):
    """[stats.truncnorm.rvs(0, 1, loc=m, scale=0.2, size=n).tolist() for m in means] (local_search.py:140-146) in one
    vectorised step.  scipy's rvs is `ppf(random_state.uniform(size), a, b) * scale + loc` on numpy's global stream
    (0.27 ms of distribution infrastructure per call, most of a hill-climb step on the host); the same three steps are
    done here for all numeric parameters of a configuration at once: the uniforms are drawn in the same order
    (row-major), ppf is elementwise, so values and stream position are identical (checked against the reference's
    neighbour sets in tests/golden/neighbors.npz)."""
    ppf = getattr(stats.truncnorm, "_ppf", None)
    if ppf is None or len(means) == 0:
        return [stats.truncnorm.rvs(0, 1, loc=m, scale=0.2, size=n).tolist() for m in means]
    u = np.random.uniform(size=(len(means), n))
    z = ppf(u.ravel(), 0.0, 1.0).reshape(len(means), n)
    return [(z[r] * 0.2 + m).tolist() for r, m in enumerate(means)]


def get_neighbors(configuration, param_space):
    """SMAC-style one-exchange neighbourhood (local_search.py:83-161), same enumeration order, same RNG draws."""
    neighbors = []
    objects = param_space.get_input_parameters_objects()
    inputs = param_space.get_input_parameters()
    numeric_neighbors = 4
    base = [configuration[p] for p in inputs]

    def push(pos, value):
        cand = list(base)
        cand[pos] = value
        if cand not in neighbors:
            neighbors.append(cand)

    # the truncated-normal draws of all real / integer parameters, in parameter order (= the reference's RNG order)
    numeric = [name for name in inputs if param_space.get_type(name) in ("real", "integer")]
    means = {}
    for name in numeric:
        lo, hi = objects[name].get_min(), objects[name].get_max()
        means[name] = (configuration[name] - lo) / (hi - lo)
    drawn = dict(zip(numeric, _truncnorm_draws([means[name] for name in numeric], numeric_neighbors)))

    for pos, name in enumerate(inputs):
        kind = param_space.get_type(name)
        obj = objects[name]
        if kind == "categorical" or (kind == "ordinal" and len(obj.get_discrete_values()) <= 5):
            for value in obj.get_discrete_values():
                push(pos, value)
        elif kind == "ordinal":
            values = obj.get_discrete_values()
            idx = values.index(configuration[name])
            if idx < numeric_neighbors / 2:
                window = values[: numeric_neighbors + 1]
            elif (len(values) - idx) <= numeric_neighbors / 2:
                window = values[-(numeric_neighbors + 1):]
            else:
                window = values[int(idx - numeric_neighbors / 2): int(idx + numeric_neighbors / 2 + 1)]
            for value in window:
                push(pos, value)
        elif kind in ("real", "integer"):
            mean = means[name]
            draws = drawn[name]
            draws.append(mean)
            for value in draws:
                value = 0 if value < 0 else (1 if value > 1 else value)
                push(pos, obj.from_range_0_1_to_parameter_value(value))
        else:
            print("Unsupported parameter type:", kind)
            raise SystemExit
    return neighbors


def _device_scorer(optimization_function):
    """Acquisition plug-ins of this package carry `device_scores(candidates, **params) -> CUDA tensor`
    (run_acquisition_function, pibo.prior_weighted_acquisition_function,
    prior_optimization.compute_EI_from_posteriors): pools are scored as arrays and stay on the GPU."""
    return getattr(optimization_function, "device_scores", None)


def _score_pool(pool, inputs, optimization_function, params, device_scorer):
    """pool: object/float array [N][D] (or list of dicts when the whole space was enumerated).
    Returns (values, n_evaluated) where values is a CUDA tensor (device scorer) or a host list."""
    if device_scorer is not None:
        if hasattr(pool, "candidates"):
            cand = pool.candidates
        else:
            cand = pool if isinstance(pool, list) else np.asarray(pool, dtype=np.float64)
        scores = device_scorer(cand, **params)
        return scores, scores.numel()
    configs = pool if isinstance(pool, list) else _rows_to_dicts(pool, inputs, range(len(pool)))
    values, _ = optimization_function(configurations=configs, **params)
    return list(values), len(values)


def _rows_to_dicts(pool, inputs, rows):
    if hasattr(pool, "candidates"):
        return pool.rows(rows)
    if isinstance(pool, list):
        return [pool[r] for r in rows]
    return [{k: pool[r, j] for j, k in enumerate(inputs)} for r in rows]


def _values_at(values, rows):
    if isinstance(values, list):
        return [values[r] for r in rows]
    if len(rows) == 0:
        return []
    import torch
    if isinstance(rows, list) and len(rows) == values.numel() and rows[0] == 0 and rows[-1] == len(rows) - 1:
        return values.cpu().tolist()                     # everything, in order
    return values[torch.as_tensor(np.asarray(rows, dtype=np.int64), device=values.device)].cpu().tolist()


def _exact_rows_fn(exact_fn, params, rows_of_positions):
    """positions -> exact reference values, or None when the scorer has no exact re-score"""
    if exact_fn is None:
        return None

    def f(positions):
        return exact_fn(np.asarray(rows_of_positions(positions), dtype=np.float64), **params)
    return f


def _hill_climb(start, param_space, inputs, optimization_function, params, enable_feasible_predictor,
                scalarization_key, feasible_parameter, exact_fn=None):
    """One start of the multi-start local search (local_search.py:337-386)."""
    iteration_data = {}
    configuration = start
    log = "Starting local search on configuration: " + str(configuration) + "\n"
    while configuration is not None:
        neighbors = data_tuples_to_dict_list(get_neighbors(configuration, param_space), inputs)
        values, feas = optimization_function(configurations=neighbors, **params)
        m = len(values)
        batch = concatenate_list_of_dictionaries(neighbors[:m])
        batch[scalarization_key] = values
        if enable_feasible_predictor:
            batch[feasible_parameter] = feas
        iteration_data = concatenate_data_dictionaries(iteration_data, batch)
        if enable_feasible_predictor:
            best = get_min_feasible_configurations(batch, 1, scalarization_key, feasible_parameter)
        elif exact_fn is not None and m > 0:
            q = exact_argmin_host(list(values), _exact_rows_fn(
                exact_fn, params, lambda pos: [[float(neighbors[i][p]) for p in inputs] for i in pos]))
            best = {key: [col[q]] for key, col in batch.items()}
        else:
            best = get_min_configurations(batch, 1, scalarization_key)
        current = {k: [v] for k, v in configuration.items()}
        if are_configurations_equal(best, current, inputs):
            log += "Local minimum found: " + str(dict(best)) + "\n"
            configuration = None
        else:
            log += "Replacing configuration by best neighbor: " + str(dict(best)) + "\n"
            configuration = {k: v[0] for k, v in best.items()}
    return iteration_data, log


def _neighbors_use_rng(param_space):
    """get_neighbors draws truncated normals only for real / integer parameters (local_search.py:138-157)."""
    return any(param_space.get_type(p) in ("real", "integer") for p in param_space.get_input_parameters())


# Lock-step hill climb: every start advances one step per round and the neighbour batches of
# all live starts are scored in ONE launch.  The climbs are independent, so the trajectories -- and the returned
# data_array, which is assembled per start afterwards -- are identical to running the starts one after the other,
# PROVIDED get_neighbors consumes no random numbers (discrete spaces).  With real / integer parameters the order of the
# truncnorm draws would change, so those spaces keep the sequential order unless set_lockstep(True) is called.
_lockstep = os.environ.get("HM_B200_LOCKSTEP", "auto")


# Neighbourhoods and moves on the GPU (hm_ls_neighbors / hm_ls_move) when the lock-step climb runs on a discrete space
_device_climb = os.environ.get("HM_B200_DEVICE_CLIMB", "1") not in ("", "0")


def set_device_climb(enabled):
    global _device_climb
    _device_climb = bool(enabled)


def _all_discrete(param_space):
    return all(param_space.get_type(p) in ("ordinal", "categorical") for p in param_space.get_input_parameters())


def set_lockstep(mode):
    """True / False / "auto" (default: lock-step only when it is exactly equivalent)."""
    global _lockstep
    _lockstep = mode


def _hill_climb_lockstep(start_list, param_space, inputs, device_scorer, params, scalarization_key, exact_fn=None):
    """All starts in lock-step; returns [(iteration_data, log)] in start order.  Rows and scores are kept as plain
    lists per start and turned into the reference's dict-of-lists once, at the end (the per-round dict concatenations
    of the sequential version are most of its host time)."""
    n = len(start_list)
    current = [[c[p] for p in inputs] for c in start_list]
    rows_of = [[] for _ in range(n)]
    vals_of = [[] for _ in range(n)]
    logs = ["Starting local search on configuration: " + str(c) + "\n" for c in start_list]
    live = list(range(n))
    while live:
        batches = [get_neighbors(dict(zip(inputs, current[s])), param_space) for s in live]
        flat = np.asarray([row for b in batches for row in b], dtype=np.float64)
        scores = device_scorer(flat, **params).cpu().tolist()
        nxt = []
        pos = 0
        for s, b in zip(live, batches):
            values = scores[pos:pos + len(b)]
            pos += len(b)
            rows_of[s].extend(b)
            vals_of[s].extend(values)
            # get_min_configurations(batch, 1): first index of the minimum, NaN first, -0.0 == +0.0 (near ties between
            # distinct values are settled with the reference's arithmetic)
            k = exact_argmin_host(values, _exact_rows_fn(exact_fn, params, lambda pos, b=b: [b[i] for i in pos]))
            best = dict(zip(inputs, b[k]))
            best[scalarization_key] = values[k]
            shown = {key: [v] for key, v in best.items()}
            if b[k] == current[s]:
                logs[s] += "Local minimum found: " + str(shown) + "\n"
            else:
                logs[s] += "Replacing configuration by best neighbor: " + str(shown) + "\n"
                current[s] = list(b[k])
                nxt.append(s)
        live = nxt
    out = []
    for s in range(n):
        data = {p: [row[j] for row in rows_of[s]] for j, p in enumerate(inputs)}
        data[scalarization_key] = vals_of[s]
        out.append((data, logs[s]))
    return out


def neighbor_plan(param_space):
    """For a space without real / integer parameters: the enumeration get_neighbors follows, as K x (parameter,
    rank inside that parameter's emitted list).  Every configuration has exactly K neighbours there: all values of a
    categorical / small ordinal, the 5-window of a larger ordinal, minus -- for every parameter but the first -- the
    entry that repeats the unchanged configuration (local_search.py:96-137)."""
    inputs = param_space.get_input_parameters()
    objs = param_space.get_input_parameters_objects()
    plan = []
    for j, name in enumerate(inputs):
        kind = param_space.get_type(name)
        n = len(objs[name].get_discrete_values())
        count = n if (kind == "categorical" or n <= 5) else 5
        if j > 0:
            count -= 1
        plan += [(j, e) for e in range(count)]
    return np.asarray(plan, dtype=np.int32).reshape(-1, 2)


def _hill_climb_device(start_list, param_space, inputs, device_scorer, params, scalarization_key, exact_fn=None):
    """The lock-step hill climb with the neighbourhoods generated and the moves decided on the GPU (hm_ls_neighbors /
    hm_ls_move); per round: one neighbour launch, the scoring launches, one move launch and one 4-byte-per-start
    read-back of the convergence flags.  Same trajectories, same returned data as _hill_climb_lockstep."""
    import ctypes
    from . import _lib
    from .space_device import device_space_for, P_ORDINAL
    torch = _lib.require_cuda()
    L = _lib.lib()
    ds = device_space_for(param_space)
    D = ds.D
    n = len(start_list)
    plan = neighbor_plan(param_space)
    K = len(plan)
    plan_d = torch.from_numpy(plan).cuda()
    current = torch.from_numpy(np.array([[float(c[p]) for p in inputs] for c in start_list], dtype=np.float64)
                               .reshape(n, D)).cuda()
    converged = torch.zeros((n,), dtype=torch.int32, device="cuda")
    best_q = torch.zeros((n,), dtype=torch.int32, device="cuda")
    live = list(range(n))
    rounds = []                      # (live list, rows [n_live*K][D], scores [n_live*K], best_q snapshot, converged snapshot)
    while live:
        live_d = torch.tensor(live, dtype=torch.int32, device="cuda")
        start_of_round = current.clone() if exact_fn is not None else None
        rows = torch.empty((len(live) * K, D), dtype=torch.float64, device="cuda")
        _lib.check(L.hm_ls_neighbors(ds.handle, _lib.dptr(plan_d), K, _lib.dptr(current), _lib.dptr(live_d), len(live),
                                     _lib.dptr(rows), _lib.stream_ptr()), "hm_ls_neighbors")
        scores = device_scorer(rows, **params).contiguous()
        _lib.check(L.hm_ls_move(_lib.dptr(scores), _lib.dptr(rows), K, D, _lib.dptr(live_d), len(live),
                                _lib.dptr(current), _lib.dptr(converged), _lib.dptr(best_q), _lib.stream_ptr()),
                   "hm_ls_move")
        if exact_fn is not None:
            # near ties between DISTINCT values inside a neighbourhood: the device move stands unless the reference's own
            # arithmetic orders them differently (rare; the start's state is then rewritten from the host)
            sh = scores.view(len(live), K).cpu().numpy()
            with np.errstate(all="ignore"):
                mn = sh.min(axis=1, keepdims=True)
                suspicious = np.isfinite(mn[:, 0]) & ((sh <= mn + NEAR_TIE_REL * np.abs(mn)) & (sh != mn)).any(axis=1)
            for l in np.nonzero(suspicious)[0]:
                s_id = live[l]
                rows_l = rows.view(len(live), K, D)[l]
                q = exact_argmin_host(sh[l].tolist(), _exact_rows_fn(
                    exact_fn, params, lambda pos, r=rows_l: r[torch.as_tensor(pos, device="cuda")].cpu().numpy()))
                if q != int(best_q[s_id].item()):
                    best_q[s_id] = q
                    same = bool(torch.equal(rows_l[q], start_of_round[s_id]))
                    converged[s_id] = 1 if same else 0
                    current[s_id] = rows_l[q] if not same else start_of_round[s_id]
        flags = converged.cpu().numpy()
        rounds.append((live, rows, scores, best_q.cpu().numpy().copy(), flags.copy()))
        live = [s for s in live if not flags[s]]
    # ---- back to the reference's dict-of-lists, with the parameter objects' own value types ----------------------
    objs = param_space.get_input_parameters_objects()
    ordinal_values = {}
    for j, p in enumerate(inputs):
        if ds.descs[j].kind == P_ORDINAL:
            vals = objs[p].get_values()
            ordinal_values[j] = (np.asarray(vals, dtype=np.float64), vals)
    cols_of = [[[] for _ in range(D)] for _ in range(n)]
    vals_of = [[] for _ in range(n)]
    logs = ["Starting local search on configuration: " + str(c) + "\n" for c in start_list]
    for live_r, rows, scores, bq, flags in rounds:
        rows_h = rows.cpu().numpy().reshape(len(live_r), K, D)
        scores_h = scores.cpu().numpy().reshape(len(live_r), K)
        typed = []
        for j in range(D):
            col = rows_h[:, :, j]
            if j in ordinal_values:
                sv, vals = ordinal_values[j]
                idx = np.searchsorted(sv, col)
                typed.append([[vals[i] for i in r] for r in idx])
            else:
                typed.append(col.astype(np.int64).tolist())
        for l, s in enumerate(live_r):
            for j in range(D):
                cols_of[s][j].extend(typed[j][l])
            v = scores_h[l].tolist()
            vals_of[s].extend(v)
            q = int(bq[s])
            shown = {p: [typed[j][l][q]] for j, p in enumerate(inputs)}
            shown[scalarization_key] = [v[q]]
            logs[s] += ("Local minimum found: " if flags[s] else "Replacing configuration by best neighbor: ") + \
                str(shown) + "\n"
    out = []
    for s in range(n):
        data = {p: cols_of[s][j] for j, p in enumerate(inputs)}
        data[scalarization_key] = vals_of[s]
        out.append((data, logs[s]))
    return out


def local_search(local_search_starting_points, local_search_random_points, param_space,
                 fast_addressing_of_data_array, enable_feasible_predictor, optimization_function,
                 optimization_function_parameters, scalarization_key, number_of_cpus, previous_points=None,
                 profiling=None, noise=False):
    """Random pools + multi-start hill climb on the acquisition function; returns (data_array, best_configuration).
    Mirrors local_search.py:389-833 (single-process branch)."""
    t0 = datetime.datetime.now()
    inputs = param_space.get_input_parameters()
    feasible_parameter = param_space.get_feasible_parameter()[0]
    params = optimization_function_parameters
    device_scorer = None if enable_feasible_predictor else _device_scorer(optimization_function)
    exact_fn = None if enable_feasible_predictor else getattr(optimization_function, "exact_scores", None)
    if exact_fn is not None and exact_fn(np.zeros((0, len(inputs))), **params) is None:
        exact_fn = None                                       # e.g. GP models: no bit-exact host restatement
    oversampling_factor = 2

    # (kept for RNG / side-effect parity with local_search.py:429-432; the hash table copy is not used afterwards)
    default_configuration = param_space.get_default_or_random_configuration()
    param_space.get_unique_hash_string_from_values(default_configuration)

    if param_space.get_space_size() < local_search_random_points:
        everything = dict_of_lists_to_list_of_dicts(param_space.get_space())
        half = int(len(everything) / 2)
        pool_u, pool_p = everything[0:half], everything[half::]
    elif device_scorer is not None and _use_device_pools(local_search_random_points):
        from .space_device import DevicePool, device_space_for
        seed = int(np.random.randint(0, 2 ** 62))            # reproducible under np.random.seed
        want64 = params.get("model_type") == "gaussian_process"
        ds = device_space_for(param_space)
        pool_u = DevicePool(ds, seed, False, local_search_random_points, want64=want64)
        pool_p = DevicePool(ds, seed, True, local_search_random_points, want64=want64)
    else:
        pool_u = param_space.get_random_configuration(size=local_search_random_points, use_priors=False, return_as_array=True)
        pool_p = param_space.get_random_configuration(size=local_search_random_points, use_priors=True, return_as_array=True)
    sampling_time = datetime.datetime.now()
    _log("Total RS time %10.4f sec\n" % ((sampling_time - t0).total_seconds()))

    # ---- hot loop A: score both pools --------------------------------------------------------------------
    val_u, n_u = _score_pool(pool_u, inputs, optimization_function, params, device_scorer)
    val_p, n_p = _score_pool(pool_p, inputs, optimization_function, params, device_scorer)
    acquisition_time = datetime.datetime.now()
    _log("Optimization function time %10.4f sec\n" % (acquisition_time - sampling_time).total_seconds())
    end_of_search = (n_u < len(pool_u)) or (n_p < len(pool_p))
    if end_of_search:
        _log("Out of budget, not all configurations were evaluated, stopping local search\n")

    # ---- seeds: 2*starting_points best of each pool, ranked separately (local_search.py:625-645) ------------
    k = local_search_starting_points * oversampling_factor

    def _select(pool, vals, kk):
        """indices of the kk smallest values in the reference's order (device top-k; near ties re-scored exactly)"""
        if exact_fn is None or isinstance(vals, list):
            return min_indices(vals, kk)
        return exact_min_indices(vals, kk, lambda idx: exact_fn(_pool_rows_matrix(pool, inputs, idx), **params))

    def best_of(pool, values, n_eval):
        if enable_feasible_predictor:
            d = concatenate_list_of_dictionaries(_rows_to_dicts(pool, inputs, range(n_eval)))
            d[scalarization_key] = list(values)
            # the plug-in reports every point feasible (random_scalarizations.py:154-155)
            d[feasible_parameter] = [1] * n_eval
            return get_min_feasible_configurations(d, k, scalarization_key, feasible_parameter)
        vals = values if not isinstance(values, list) else values[:n_eval]
        rows = [int(i) for i in _select(pool, vals, k)]
        d = concatenate_list_of_dictionaries(_rows_to_dicts(pool, inputs, rows)) if rows else {p: [] for p in inputs}
        d[scalarization_key] = _values_at(values, rows)
        return d

    seeds_u = best_of(pool_u, val_u, n_u)
    seeds_p = best_of(pool_p, val_p, n_p)
    keys = inputs + [scalarization_key]
    seeds = {key: list(seeds_u[key]) + list(seeds_p[key]) for key in keys}
    n_seed_u, n_seed_p = len(seeds_u[scalarization_key]), len(seeds_p[scalarization_key])

    best_previous = None
    if previous_points is not None:
        if enable_feasible_predictor:
            best_previous = get_min_feasible_configurations(previous_points, local_search_starting_points,
                                                            scalarization_key, feasible_parameter)
        else:
            best_previous = get_min_configurations(previous_points, local_search_starting_points, scalarization_key)
        seeds = {key: seeds[key] + list(best_previous[key]) for key in keys}

    # duplicate filter (local_search.py:673-699 -> space.remove_duplicate_configs): first occurrence wins in the order
    # previous > prior > uniform; every group comes back sorted lexicographically by parameter values.
    table_keys = keys
    if _dedup_column_order is not None:
        if sorted(_dedup_column_order) != sorted(keys):
            raise ValueError("set_dedup_column_order: keys must be a permutation of %r" % (keys,))
        table_keys = list(_dedup_column_order)
    table = np.zeros((len(seeds[scalarization_key]), len(table_keys)), dtype=np.dtype(object))
    for j, key in enumerate(table_keys):
        table[:, j] = seeds[key]
    grp_u = table[0:n_seed_u]
    grp_p = table[n_seed_u:n_seed_u + n_seed_p]
    grp_prev = table[n_seed_u + n_seed_p:]
    grp_prev, grp_p, grp_u = param_space.remove_duplicate_configs(grp_prev, grp_p, grp_u,
                                                                  ignore_columns=table_keys.index(scalarization_key))
    chosen = np.concatenate((grp_u[0:local_search_starting_points], grp_p[0:local_search_starting_points],
                             grp_prev[0:local_search_starting_points]), axis=0)
    starts = {key: chosen[:, j].tolist() for j, key in enumerate(table_keys)}
    data_collection_time = datetime.datetime.now()
    n_starts = len(starts[scalarization_key])
    _log("Starting local search iteration: " + ", #configs:" + str(n_starts) + "\n")

    # ---- hot loop B: multi-start hill climb, starts in order (keeps the RNG order of get_neighbors) ----------
    result_array = {}
    start_list = [{key: starts[key][s] for key in keys} for s in range(n_starts)]
    # "auto": only when it is exactly equivalent -- no RNG in get_neighbors, and a scorer whose value for a candidate
    # does not depend on what else is in the batch (piBO's TS/UCB branch and BOPrO's running limits do)
    lockstep = device_scorer is not None and (_lockstep is True or _lockstep in ("1", "on") or
                                              (_lockstep == "auto" and not _neighbors_use_rng(param_space) and
                                               getattr(device_scorer, "batch_invariant", False)))
    if lockstep and _device_climb and _all_discrete(param_space) and n_starts > 0:
        climbs = _hill_climb_device(start_list, param_space, inputs, device_scorer, params, scalarization_key, exact_fn)
    elif lockstep:
        climbs = _hill_climb_lockstep(start_list, param_space, inputs, device_scorer, params, scalarization_key, exact_fn)
    else:
        climbs = (_hill_climb(start, param_space, inputs, optimization_function, params, enable_feasible_predictor,
                              scalarization_key, feasible_parameter, exact_fn) for start in start_list)
    for it_data, log in climbs:
        _log(log, verbose=True)
        result_array = concatenate_data_dictionaries(result_array, it_data)
    local_search_time = datetime.datetime.now()
    _log("Multi-start LS time %10.4f sec\n" % (local_search_time - acquisition_time).total_seconds())

    # ---- final pick: argmin over [LS results, uniform pool, prior pool, previous points] skipping configurations
    #      that were already evaluated (local_search.py:789-807); ties -> first in that order -------------------
    ls_vals = list(result_array.get(scalarization_key, []))
    prev_vals = list(previous_points[scalarization_key]) if previous_points is not None else []
    off_u = len(ls_vals)
    off_p = off_u + n_u
    off_prev = off_p + n_p

    def config_at(pos):
        if pos < off_u:
            return {p: result_array[p][pos] for p in inputs}
        if pos < off_p:
            return _rows_to_dicts(pool_u, inputs, [pos - off_u])[0]
        if pos < off_prev:
            return _rows_to_dicts(pool_p, inputs, [pos - off_p])[0]
        return {p: previous_points[p][pos - off_prev] for p in inputs}

    def rank_key(v, pos):
        return (0, 0.0, pos) if v != v else (1, v + 0.0, pos)      # NaN first, -0.0 == +0.0, then position

    best_configuration = None
    want = 32
    while best_configuration is None:
        cand = [(v, i) for i, v in enumerate(ls_vals)]
        horizons = []
        for pool, values, n_eval, off in ((pool_u, val_u, n_u, off_u), (pool_p, val_p, n_p, off_p)):
            vals = values if not isinstance(values, list) else values[:n_eval]
            rows = [int(i) for i in _select(pool, vals, min(want, n_eval))]
            fetched = [(v, off + r) for v, r in zip(_values_at(values, rows), rows)]
            cand += fetched
            if n_eval > want and fetched:
                horizons.append(rank_key(*fetched[-1]))     # unfetched entries of this pool rank after this pair
        cand += [(v, off_prev + i) for i, v in enumerate(prev_vals)]
        cand.sort(key=lambda c: rank_key(*c))
        need_more = False
        eligible = []          # the first configuration not evaluated yet + the ones whose value is a near tie with it
        limit = None
        for v, pos in cand:
            if limit is not None and not (v <= limit):
                break
            if any(rank_key(v, pos) > h for h in horizons):
                need_more = True
                break
            conf = config_at(pos)
            if param_space.get_unique_hash_string_from_values(conf) not in fast_addressing_of_data_array:
                eligible.append((v, pos, conf))
                if limit is None:
                    if exact_fn is None or v != v or v in (float("inf"), float("-inf")):
                        break
                    limit = v + NEAR_TIE_REL * abs(v)
        if need_more:
            want *= 8                                       # the answer (or its near-tie window) lies beyond what was fetched
            continue
        if not eligible:
            raise ValueError("attempt to get argmin of an empty sequence")   # what np.argmin raises in the reference
        pick = 0
        if len({v + 0.0 for v, _, _ in eligible}) > 1:
            # distinct values closer than 1e-9 relative: order them with the reference's own arithmetic
            exact = exact_fn(np.array([[float(c[p]) for p in inputs] for _, _, c in eligible], dtype=np.float64), **params)
            pick = min(range(len(eligible)), key=lambda i: rank_key(float(exact[i]), eligible[i][1]))
        best_configuration = eligible[pick][2]
    post_time = datetime.datetime.now()
    _log("MSLS time %10.4f sec\n" % (post_time - acquisition_time).total_seconds())

    # ---- returned data_array: [LS results, pools, previous] as dict of lists (pools only when small) ---------
    data_array = {}
    if n_u + n_p <= MATERIALISE_LIMIT:
        def columns(pool, n):
            if isinstance(pool, np.ndarray):           # column slices of the sample array: no per-row dicts
                return {p: pool[:n, j].tolist() for j, p in enumerate(inputs)}
            return concatenate_list_of_dictionaries(_rows_to_dicts(pool, inputs, range(n))) if n else {p: [] for p in inputs}

        du = columns(pool_u, n_u)
        du[scalarization_key] = _values_at(val_u, list(range(n_u)))
        dp = columns(pool_p, n_p)
        dp[scalarization_key] = _values_at(val_p, list(range(n_p)))
        data_array = {key: list(du[key]) + list(dp[key]) for key in keys}
    else:
        data_array = {key: list(seeds_u[key]) + list(seeds_p[key]) for key in keys}
    if previous_points is not None:
        data_array = {key: data_array[key] + list(previous_points[key]) for key in keys}
    if result_array:
        data_array = {key: list(result_array[key]) + data_array[key] for key in keys}

    if profiling is not None:
        profiling.add("(LS) Random sampling time", (sampling_time - t0).total_seconds())
        profiling.add("(LS) Acquisition evaluation time", (acquisition_time - sampling_time).total_seconds())
        profiling.add("(LS) Data collection time", (data_collection_time - acquisition_time).total_seconds())
        profiling.add("(LS) Multi-start LS time", (local_search_time - data_collection_time).total_seconds())
        profiling.add("(LS) Post-MSLS data treatment time", (post_time - local_search_time).total_seconds())
    return data_array, best_configuration
