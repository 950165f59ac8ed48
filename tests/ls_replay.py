"""Replay of the reference's local_search() runs frozen in tests/golden/local_search_traj.npz (make_golden.py
local_search_traj_case): same fitted forests, the two random pools the reference drew, and numpy's RNG state right
after the pools (the neighbourhoods of real / integer parameters consume it, local_search.py:138-157)."""
import copy
import json
import os
import random

import numpy as np

from conftest import GOLDEN

TAGS = ["branin_s0", "branin_s1", "branin_s2", "mixed_s0", "mixed_s1", "mixed_s2",
        "discrete_s0", "discrete_s1", "discrete_s2", "discrete_s3"]


def load():
    return np.load(os.path.join(GOLDEN, "local_search_traj.npz"), allow_pickle=False)


def typed_column(param, col):
    """float column of the golden file -> the python types the reference's sample arrays hold"""
    if param["type"] in ("integer", "categorical"):
        return [int(v) for v in col]
    if param["type"] == "ordinal" and all(float(x).is_integer() for x in param["values"]):
        return [int(v) for v in col]
    return [float(v) for v in col]


def replay(z, tag, optimization_function, frozen_forest_cls):
    """Runs hypermapper_b200.local_search.local_search on the frozen inputs of run `tag`.
    Returns (data_array, best, golden data_array [D+1][n], golden best, inputs, scalarization key)."""
    from hypermapper_b200 import local_search as ls, utility_functions as uf
    from bench_support.space_lite import SpaceLite
    name = tag.rsplit("_", 1)[0]
    config = json.loads(str(z[name + "_config_json"]))
    params_desc = json.loads(str(z[name + "_params_json"]))
    meta = json.loads(str(z[tag + "_meta"]))
    space = SpaceLite(config)
    inputs = space.get_input_parameters()
    objectives = config["optimization_objectives"]
    models = {o: frozen_forest_cls(z, "%s_forest_%s_" % (tag, o)) for o in objectives}
    di = z[tag + "_data_inputs"]
    data = {p["name"]: typed_column(p, di[:, j]) for j, p in enumerate(params_desc)}
    for o in objectives:
        data[o] = z["%s_data_%s" % (tag, o)].tolist()
    n_train = len(data[objectives[0]])
    fast = {space.get_unique_hash_string_from_values({k: data[k][r] for k in inputs}): r for r in range(n_train)}
    weights = {o: float(w) for o, w in zip(objectives, z[tag + "_weights"])}
    limits = {o: [min(data[o]), max(data[o])] for o in objectives}
    key = config["scalarization_key"]
    d2 = copy.deepcopy(data)
    s, _ = uf.compute_data_array_scalarization(d2, weights, copy.deepcopy(limits), meta["scal"])
    d2[key] = s.tolist()
    p = {"regression_models": models, "iteration_number": meta["iteration"], "data_array": d2,
         "classification_model": None, "param_space": space, "objective_weights": weights,
         "model_type": "random_forest", "objective_limits": copy.deepcopy(limits),
         "acquisition_function": meta["acq"], "scalarization_method": meta["scal"], "number_of_cpus": 1}
    pools = [z[tag + "_pool_uniform"], z[tag + "_pool_prior"]]
    state = ("MT19937", z[tag + "_rng_keys"], int(z[tag + "_rng_pos"][0]), int(z[tag + "_rng_pos"][1]),
             float(z[tag + "_rng_gauss"][0]))

    def frozen_pools(use_priors=True, size=1, return_as_array=False):
        arr = pools.pop(0)
        out = np.empty(arr.shape, dtype=object)
        for j, pd in enumerate(params_desc):
            out[:, j] = typed_column(pd, arr[:, j])
        if not pools:
            np.random.set_state(state)          # the stream position the reference had after drawing both pools
        return out

    space.get_random_configuration = frozen_pools
    ls.set_dedup_column_order(json.loads(str(z[tag + "_dedup_columns"])))
    np.random.seed(meta["seed"])
    random.seed(meta["seed"])
    try:
        data_array, best = ls.local_search(meta["n_starts"], meta["n_points"], space, fast, False, optimization_function,
                                           p, key, 1, previous_points=d2, profiling=None)
    finally:
        ls.set_dedup_column_order(None)
    return data_array, best, z[tag + "_data_array"], z[tag + "_best"], inputs, key


def check(data_array, best, gold_data, gold_best, inputs, key, rel):
    assert [float(best[k]) for k in inputs] == gold_best.tolist()
    got = np.array([[float(v) for v in data_array[k]] for k in inputs + [key]], dtype=np.float64)
    assert got.shape == gold_data.shape
    assert np.array_equal(got[:-1], gold_data[:-1])                    # every evaluated configuration, in order
    a, b = got[-1], gold_data[-1]
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    err[a == b] = 0
    assert np.all(err <= rel), float(err.max())
