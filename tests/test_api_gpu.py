"""GPU parity through the reference-facing Python API (the drop-in boundary): the functions named like the
reference's, called the way the reference calls them, against golden vectors produced by the reference itself."""
import copy
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

REL = 1e-5


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


class FrozenForest:
    """A fitted model as the exporter sees it (estimators_ with a tree_, Hutter tables), rebuilt from the golden file."""

    class _Tree:
        def __init__(self, left, right, feature, threshold, value):
            self.children_left, self.children_right, self.feature, self.threshold = left, right, feature, threshold
            self.value = value.reshape(-1, 1, 1)
            self.node_count = len(left)

    class _Est:
        def __init__(self, tree):
            self.tree_ = tree

    def __init__(self, z, prefix):
        off = z[prefix + "tree_off"]
        self.estimators_ = []
        self.tree_means_per_leaf, self.tree_vars_per_leaf = [], []
        for t in range(len(off) - 1):
            b, e = int(off[t]), int(off[t + 1])
            left = z[prefix + "left"][b:e]
            self.estimators_.append(self._Est(self._Tree(left, z[prefix + "right"][b:e], z[prefix + "feature"][b:e],
                                                         z[prefix + "threshold"][b:e], z[prefix + "value"][b:e])))
            leaves = np.nonzero(left < 0)[0]
            self.tree_means_per_leaf.append({int(l): float(z[prefix + "leaf_mean"][b + l]) for l in leaves})
            self.tree_vars_per_leaf.append({int(l): float(z[prefix + "leaf_var"][b + l]) for l in leaves})
        self.n_features_in_ = int(z[prefix + "n_features"])


def setup_case(name):
    from bench_support.space_lite import SpaceLite
    z = load(name)
    config = json.loads(str(z["config_json"]))
    space = SpaceLite(config)
    objectives = json.loads(str(z["objectives_json"]))
    params = json.loads(str(z["params_json"]))
    models = {o: FrozenForest(z, "forest_%s_" % o) for o in objectives}
    cand = z["candidates"]
    bufferx = []
    for row in cand:
        t = []
        for p, v in zip(params, row):
            if p["type"] in ("integer", "categorical"):
                t.append(int(v))
            elif p["type"] == "ordinal" and all(float(x).is_integer() for x in p["values"]):
                t.append(int(v))
            else:
                t.append(float(v))
        bufferx.append(tuple(t))
    data = {o: z["data_" + o].tolist() for o in objectives}
    weights = {o: float(w) for o, w in zip(objectives, z["weights"])}
    limits = {o: [float(a), float(b)] for o, (a, b) in zip(objectives, z["stored_limits"])}
    return z, space, objectives, models, bufferx, data, weights, limits


def rel_ok(a, b, rel):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    err[a == b] = 0
    return np.all(err <= rel)


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
def test_preprocess_and_mean_uncertainty(name):
    from hypermapper_b200 import models as hm
    z, space, objectives, models, bufferx, *_ = setup_case(name)
    enc = hm.preprocess_data_buffer(bufferx, space)
    assert np.array_equal(np.array(enc), z["encoded"])
    mean, var = hm.compute_model_mean_and_uncertainty(bufferx, models, "random_forest", space, var=True)
    _, std = hm.compute_model_mean_and_uncertainty(bufferx, models, "random_forest", space, var=False)
    assert list(mean.keys()) == objectives
    for o in objectives:
        assert np.array_equal(mean[o], z["mean_" + o])
        assert rel_ok(var[o], z["var_" + o], 1e-10)
        assert rel_ok(std[o], z["std_" + o], 1e-10)
        leaves = hm.rf_get_leaves_per_sample(models[o], z["encoded"])
        assert leaves.dtype == np.int64 and np.array_equal(leaves, z["leaves_" + o])
        assert np.array_equal(hm.rf_compute_rf_prediction(models[o], leaves), z["mean_" + o])
        assert np.array_equal(hm.rf_compute_rf_prediction_variance(models[o], leaves, z["mean_" + o]), z["var_" + o])
    pred = hm.model_prediction(bufferx, models, space)
    for o in objectives:
        assert np.array_equal(pred[o], z["sk_predict_" + o])
    # 2-D array input takes the same path
    mean2, _ = hm.compute_model_mean_and_uncertainty(z["candidates"], models, "random_forest", space, var=True)
    for o in objectives:
        assert np.array_equal(mean2[o], z["mean_" + o])


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
@pytest.mark.parametrize("acq", ["EI", "UCB", "TS"])
@pytest.mark.parametrize("scal", ["linear", "tchebyshev", "modified_tchebyshev"])
def test_run_acquisition_function_list_of_dicts(name, acq, scal):
    from hypermapper_b200 import random_scalarizations as rs
    z, space, objectives, models, bufferx, data, weights, limits = setup_case(name)
    inputs = space.get_input_parameters()
    configurations = [dict(zip(inputs, row)) for row in bufferx]
    lim_before = copy.deepcopy(limits)
    vals, feas = rs.run_acquisition_function(acq, configurations, weights, models, space, scal, limits,
                                             int(z["iteration_number"]), data, "random_forest", None, 1)
    assert isinstance(vals, list) and feas == [1] * len(configurations)
    assert limits == lim_before                       # inputs are not mutated
    assert rel_ok(vals, z["acq_%s_%s" % (acq, scal)], REL)


def test_unknown_enums_raise_system_exit():
    from hypermapper_b200 import random_scalarizations as rs
    z, space, objectives, models, bufferx, data, weights, limits = setup_case("rf_branin.npz")
    cfgs = [dict(zip(space.get_input_parameters(), bufferx[0]))]
    with pytest.raises(SystemExit):
        rs.run_acquisition_function("PI", cfgs, weights, models, space, "linear", limits, 1, data, "random_forest")
    with pytest.raises(SystemExit):
        rs.run_acquisition_function("EI", cfgs, weights, models, space, "chebyshev", limits, 1, data, "random_forest")


@pytest.mark.parametrize("k", [1, 20, 64])
def test_get_min_configurations_dict_api(k):
    from hypermapper_b200 import utility_functions as uf
    z = load("rf_mixed.npz")
    vals = z["acq_EI_tchebyshev"]
    d = {"__index": list(range(len(vals))), "scal": list(vals)}
    best = uf.get_min_configurations(d, k, "scal")
    assert best["__index"] == z["topk_idx_%d" % k].tolist()
    assert best["scal"] == z["topk_val_%d" % k].tolist()
    assert len(d["scal"]) == len(vals)                # input not consumed


def test_feasibility_classifier_probability_matches_sklearn():
    """N1: P(feasible) through the traversal kernel == RandomForestClassifier.predict_proba (class_weight etc. as
    models.py:574-578), and it multiplies the acquisition value."""
    from sklearn.ensemble import RandomForestClassifier
    from hypermapper_b200 import models as hm, random_scalarizations as rs
    from bench_support.space_lite import SpaceLite
    z, space, objectives, models, bufferx, data, weights, limits = setup_case("rf_mixed.npz")
    config = json.loads(str(z["config_json"]))
    config["feasible_output"] = {"enable_feasible_predictor": True, "name": "Valid", "true_value": "True", "false_value": "False"}
    space = SpaceLite(config)
    Xtr = z["encoded"][:300]
    ytr = (Xtr[:, 0] + Xtr[:, 1] > np.median(Xtr[:, 0] + Xtr[:, 1]))
    clf = RandomForestClassifier(class_weight={True: 0.9, False: 0.1}, n_estimators=10, max_features=0.75, random_state=0)
    clf.fit(Xtr, ytr)
    probs = hm.model_probabilities(bufferx, {"Valid": clf}, space)["Valid"]
    want = clf.predict_proba(z["encoded"])
    assert np.array_equal(probs, want)
    inputs = space.get_input_parameters()
    configurations = [dict(zip(inputs, row)) for row in bufferx]
    vals, _ = rs.run_acquisition_function("EI", configurations, weights, models, space, "tchebyshev", limits,
                                          int(z["iteration_number"]), data, "random_forest", {"Valid": clf}, 1)
    true_idx = clf.classes_.tolist().index(True)
    assert rel_ok(vals, z["acq_EI_tchebyshev"] * want[:, true_idx], REL)


def test_local_search_matches_reference_golden():
    """local_search() end point on the discrete golden scenario: same pools (frozen from the reference's RNG draws),
    same fitted forest -> same best configuration as the reference picked."""
    from hypermapper_b200 import random_scalarizations as rs, utility_functions as uf
    from bench_support.space_lite import SpaceLite
    z = load("local_search.npz")
    config = json.loads(str(z["config_json"]))
    space = SpaceLite(config)
    inputs = space.get_input_parameters()
    models = {"Value": FrozenForest(z, "forest_Value_")}
    data = {k: [int(v) for v in z["data_inputs"][:, j]] for j, k in enumerate(inputs)}
    data["Value"] = z["data_Value"].tolist()
    fast = {space.get_unique_hash_string_from_values({k: data[k][r] for k in inputs}): r for r in range(60)}
    weights = {"Value": 1.0}
    limits = {"Value": [min(data["Value"]), max(data["Value"])]}
    for trial, seed in enumerate((3, 4, 5)):
        for acq in ("EI", "UCB"):
            pools = [z["pool_uniform_%s_%d" % (acq, seed)], z["pool_prior_%s_%d" % (acq, seed)]]

            def frozen_pools(use_priors=True, size=1, return_as_array=False, _p=pools):
                arr = _p.pop(0)
                out = np.array([[None] * arr.shape[1]] * arr.shape[0])
                for j in range(arr.shape[1]):
                    out[:, j] = arr[:, j].astype(np.int64)
                return out

            space.get_random_configuration = frozen_pools
            d2 = copy.deepcopy(data)
            scal, _ = uf.compute_data_array_scalarization(d2, weights, copy.deepcopy(limits), "tchebyshev")
            d2[config["scalarization_key"]] = scal.tolist()
            cfg2 = copy.deepcopy(config)
            cfg2["acquisition_function"] = acq
            np.random.seed(seed)
            random.seed(seed)
            best = rs.random_scalarizations(cfg2, d2, space, fast, models, 3 + trial, weights, copy.deepcopy(limits),
                                            None, None, "local_search")
            assert [float(best[k]) for k in inputs] == z["best_%s_%d" % (acq, seed)].tolist()
            assert space.get_unique_hash_string_from_values(best) not in fast


def test_local_search_generic_callable_and_return_shape():
    """the plug-in seam: any f(configurations=[dict], **params) -> (values, feasibility) works, and the returned
    data_array holds every evaluated point (LS points first, then both pools, then previous points)."""
    from hypermapper_b200 import local_search as ls
    from bench_support.space_lite import SpaceLite
    config = {"optimization_objectives": ["v"], "input_parameters": {
        "a": {"parameter_type": "ordinal", "values": list(range(40)), "parameter_default": 0},
        "b": {"parameter_type": "integer", "values": [0, 30], "parameter_default": 0},
        "c": {"parameter_type": "real", "values": [0, 1], "parameter_default": 0.5}}}
    space = SpaceLite(config)
    calls = []

    def f(configurations, offset):
        calls.append(len(configurations))
        return [abs(c["a"] - 17) + abs(c["b"] - 4) + abs(c["c"] - 0.25) + offset for c in configurations], [1] * len(configurations)

    np.random.seed(0)
    data_array, best = ls.local_search(3, 200, space, {}, False, f, {"offset": 1.0}, "scalarization", 1)
    assert calls[0] == 200 and calls[1] == 200
    assert set(data_array.keys()) == {"a", "b", "c", "scalarization"}
    n = len(data_array["scalarization"])
    assert n == sum(calls)
    i = int(np.argmin(data_array["scalarization"]))
    assert [data_array[k][i] for k in ("a", "b", "c")] == [best[k] for k in ("a", "b", "c")]
    assert best["a"] == 17 and best["b"] == 4


def test_lockstep_hill_climb_equals_sequential():
    """The lock-step hill climb (all starts advance together, one scoring launch per round) returns exactly what the
    start-after-start climb returns on a discrete space: same data_array (order included), same pick."""
    from hypermapper_b200 import local_search as ls, random_scalarizations as rs, utility_functions as uf
    from bench_support.space_lite import SpaceLite
    z = load("local_search.npz")
    config = json.loads(str(z["config_json"]))
    space = SpaceLite(config)
    inputs = space.get_input_parameters()
    models = {"Value": FrozenForest(z, "forest_Value_")}
    data = {k: [int(v) for v in z["data_inputs"][:, j]] for j, k in enumerate(inputs)}
    data["Value"] = z["data_Value"].tolist()
    fast = {space.get_unique_hash_string_from_values({k: data[k][r] for k in inputs}): r for r in range(60)}
    weights = {"Value": 1.0}
    limits = {"Value": [min(data["Value"]), max(data["Value"])]}
    scal, _ = uf.compute_data_array_scalarization(data, weights, copy.deepcopy(limits), "tchebyshev")
    data[config["scalarization_key"]] = scal.tolist()
    p = {"regression_models": models, "iteration_number": 3, "data_array": data, "classification_model": None,
         "param_space": space, "objective_weights": weights, "model_type": "random_forest",
         "objective_limits": limits, "acquisition_function": "EI", "scalarization_method": "tchebyshev",
         "number_of_cpus": 1}
    assert not ls._neighbors_use_rng(space)
    results = []
    try:
        # sequential / lock-step with host neighbourhoods / lock-step with device neighbourhoods and moves / default
        for mode, device in ((False, False), (True, False), (True, True), ("auto", True)):
            ls.set_lockstep(mode)
            ls.set_device_climb(device)
            np.random.seed(4)
            random.seed(4)
            results.append(ls.local_search(6, 400, space, fast, False, rs.run_acquisition_function, p,
                                           config["scalarization_key"], 1, previous_points=data))
    finally:
        ls.set_lockstep("auto")
        ls.set_device_climb(True)
    for data_array, best in results[1:]:
        assert best == results[0][1]
        assert data_array == results[0][0]
        # same python types as the reference's neighbour lists (they end up in hash strings): ints stay ints
        for k in inputs:
            assert [type(v) for v in data_array[k]] == [type(v) for v in results[0][0][k]], k
    assert len(results[0][0][config["scalarization_key"]]) > 800 + 60      # the climbs did evaluate neighbours


def test_device_forest_cache_follows_in_place_refit():
    """the reference's fit_RF (models.py:184-223) calls get_leaves_per_sample in the MIDDLE of the
    fit -- before the Hutter tables exist and before tree_.threshold is rewritten in place.  Predictions made after the
    fit must see the final thresholds and tables, not a forest cached during it."""
    from hypermapper_b200 import forest, models as hm
    from bench_support.hutter_forest import HutterForest, fit_hutter_forest
    from bench_support.space_lite import SpaceLite
    from oracle import oracle
    rng = np.random.default_rng(0)
    X = rng.random((200, 3))
    y = np.sin(4 * X[:, 0]) + X[:, 1] * X[:, 2]
    final = fit_hutter_forest(X, y, n_estimators=8, seed=3)
    final_thr = [est.tree_.threshold.copy() for est in final.estimators_]
    # rewind to the state fit_RF is in when it calls get_leaves_per_sample: no tables, other thresholds
    means, variances = final.tree_means_per_leaf, final.tree_vars_per_leaf
    model = HutterForest(final.estimators_, [], [], 3)
    for est in model.estimators_:
        internal = est.tree_.children_left != est.tree_.children_right
        est.tree_.threshold[internal] += 0.05
    leaves_mid = hm.rf_get_leaves_per_sample(model, X)
    fa_mid = oracle.forest_arrays_from_model(model)
    assert np.array_equal(leaves_mid, oracle.rf_apply(fa_mid, X))
    _ = forest.device_forest(model)                       # anything cached now is stale after the next three lines
    model.tree_means_per_leaf, model.tree_vars_per_leaf = means, variances
    for est, thr in zip(model.estimators_, final_thr):
        est.tree_.threshold[:] = thr
    space = SpaceLite({"optimization_objectives": ["v"], "input_parameters": {
        "a": {"parameter_type": "real", "values": [0, 1]}, "b": {"parameter_type": "real", "values": [0, 1]},
        "c": {"parameter_type": "real", "values": [0, 1]}}})
    Xc = rng.random((500, 3))
    mean, var = hm.compute_rf_prediction_mean_and_uncertainty(Xc, {"v": model}, space, var=True)
    fa = oracle.forest_arrays_from_model(model)
    leaf = oracle.rf_apply(fa, Xc)
    want = oracle.rf_prediction(fa, leaf)
    assert np.array_equal(mean["v"], want)
    assert rel_ok(var["v"], oracle.rf_prediction_variance(fa, leaf, want), 1e-10)


def test_classifier_probability_is_normalised_per_leaf():
    """tree_.value of a classifier holds weighted class COUNTS on scikit-learn < 1.4 (the versions
    the unmodified reference runs on) and fractions later; the leaf payload must be the fraction either way."""
    from sklearn.ensemble import RandomForestClassifier
    from hypermapper_b200 import forest
    rng = np.random.default_rng(1)
    X = rng.random((300, 4))
    y = X[:, 0] + X[:, 1] > 1.0
    clf = RandomForestClassifier(n_estimators=5, random_state=0).fit(X, y)
    tabs = forest.forest_tables(forest._NoHutter(clf), value_column=1, normalize_value=True)

    class CountTree:
        def __init__(self, tr, scale):
            self.children_left, self.children_right = tr.children_left, tr.children_right
            self.feature, self.threshold, self.node_count = tr.feature, tr.threshold, tr.node_count
            self.value = tr.value * scale                     # what an old scikit-learn stores: counts, not fractions

    class Est:
        def __init__(self, tree):
            self.tree_ = tree

    class Old:
        estimators_ = [Est(CountTree(e.tree_, 37.0 + i)) for i, e in enumerate(clf.estimators_)]
        n_features_in_ = 4
        tree_means_per_leaf = tree_vars_per_leaf = None
    old = forest.forest_tables(Old, value_column=1, normalize_value=True)
    assert np.allclose(old["value"], tabs["value"], rtol=1e-15, atol=0)
    leaf = tabs["left"] < 0
    assert np.all((tabs["value"][leaf] >= 0) & (tabs["value"][leaf] <= 1))


def test_integration_md_snippet_runs():
    """INTEGRATION.md section 2 (the ctypes binding a maintainer of the reference would add) is executed as written:
    extracted from the markdown, run against the built library with a fitted model of the golden file, and compared with
    the reference's own values."""
    import re
    from conftest import ROOT
    from hypermapper_b200 import _lib
    from hypermapper_b200.random_scalarizations import acquisition_constants
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"# integration-snippet-begin\n(.*?)# integration-snippet-end", text, re.S)
    assert m, "snippet markers missing"
    z, space, objectives, models, bufferx, data, weights, limits = setup_case("rf_mixed.npz")
    w, fmin, _ = acquisition_constants("EI", "tchebyshev", objectives, weights, limits, data, int(z["iteration_number"]))
    ns = {"LIB_PATH": _lib.SO_PATH, "rf_models": models, "X_enc": z["encoded"], "weights": w, "f_min": fmin}
    exec(compile(m.group(1), "INTEGRATION.md", "exec"), ns)
    assert rel_ok(ns["scores"], z["acq_EI_tchebyshev"], REL)
    assert np.array_equal(ns["topk_idx"], z["topk_idx_20"])
