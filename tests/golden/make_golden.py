"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) in the build container.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz / *.json

The reference cannot travel to the GPU box, so its outputs are frozen here:
  rf_mixed.npz    fitted RFModel forests (raw sklearn arrays + Hutter leaf tables), candidates, and the
                  reference's own outputs: encoded buffer, leaves, mean/var/std, sklearn predict,
                  run_acquisition_function for {EI,UCB,TS} x {linear,tchebyshev,modified_tchebyshev},
                  get_min_configurations picks, compute_data_array_scalarization
  rf_branin.npz   the bundled _branin_scenario.json shape (C1): 2 real inputs, 10 trees, 1 objective
  pareto.npz      sequential_is_pareto_efficient masks on random / tie-heavy / duplicated / NaN inputs and the
                  three known-answer cases of the reference's tests (tests/data/*pareto*.csv)
  local_search.npz  local_search() trajectory end points under fixed seeds (discrete space)
  priors.npz      a space with every prior form (named Beta densities, probability lists, interpolated list prior,
                  custom Gaussian): compute_probability_from_prior, pibo.prior_weighted_acquisition_function and
                  prior_optimization.compute_EI_from_posteriors outputs of the reference
"""
import copy
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# The reference iterates Python `set`s of key strings (utility_functions.py:237-239), so some of its results (the column
# order of the duplicate filter in local_search.py:673-699, hence which seeds survive and in which order the hill climbs
# consume numpy's RNG stream) depend on the interpreter's string-hash seed.  Pin it so that this script is reproducible.
if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

from oracle import ref_harness  # noqa: E402

ns = ref_harness.load()
REF = ref_harness.REFERENCE_ROOT


def validated(config):
    from jsonschema import Draft4Validator
    with open(os.path.join(REF, "hypermapper", "schema.json")) as f:
        schema = json.load(f)
    V = ns.utility_functions.extend_with_default(Draft4Validator)
    config = copy.deepcopy(config)
    V(schema).validate(config)
    return config


MIXED = {
    "application_name": "golden_mixed",
    "optimization_objectives": ["obj_a", "obj_b"],
    "optimization_iterations": 5,
    "models": {"model": "random_forest", "number_of_trees": 8},
    "input_parameters": {
        "x_real": {"parameter_type": "real", "values": [-5, 10], "parameter_default": 0},
        "x_int": {"parameter_type": "integer", "values": [0, 20], "parameter_default": 3},
        "x_ord": {"parameter_type": "ordinal", "values": [1, 2, 4, 8, 16, 32, 64, 128], "parameter_default": 4},
        "x_cat": {"parameter_type": "categorical", "values": ["a", "b", "c"], "parameter_default": "a"},
        "x_ordf": {"parameter_type": "ordinal", "values": [0.5, 1.5, 2.5], "parameter_default": 1.5},
        "x_cat2": {"parameter_type": "categorical", "values": ["p", "q"], "parameter_default": "q"},
    },
}

BRANIN = {
    "application_name": "golden_branin",
    "optimization_objectives": ["Value"],
    "optimization_iterations": 5,
    "input_parameters": {
        "x1": {"parameter_type": "real", "values": [-5, 10], "parameter_default": 0},
        "x2": {"parameter_type": "real", "values": [0, 15], "parameter_default": 0},
    },
}

DISCRETE = {
    "application_name": "golden_discrete",
    "optimization_objectives": ["Value"],
    "optimization_iterations": 5,
    "models": {"model": "random_forest", "number_of_trees": 6},
    "local_search_random_points": 300,
    "local_search_starting_points": 4,
    "number_of_cpus": 1,
    "input_parameters": {
        "a": {"parameter_type": "ordinal", "values": list(range(1, 33)), "parameter_default": 1},
        "b": {"parameter_type": "ordinal", "values": [1, 2, 4, 8, 16, 32, 64], "parameter_default": 1},
        "c": {"parameter_type": "categorical", "values": ["u", "v", "w", "x"], "parameter_default": "u"},
        "d": {"parameter_type": "ordinal", "values": [0, 1, 2], "parameter_default": 0},
    },
}


def params_desc(config):
    out = []
    for name, p in config["input_parameters"].items():
        vals = p["values"]
        if p["parameter_type"] == "ordinal":
            vals = sorted(vals, key=float)
        out.append({"name": name, "type": p["parameter_type"], "values": vals})
    return out


def objective_mixed(cfg):
    x, i, o, c, of, c2 = cfg["x_real"], cfg["x_int"], cfg["x_ord"], cfg["x_cat"], cfg["x_ordf"], cfg["x_cat2"]
    a = (x - 2.0) ** 2 + 0.3 * i + np.log2(o) * (1 + c) + of * 2 - 3 * c2
    b = 50 - 3 * x + (i - 10) ** 2 / 4.0 + 20.0 / o + 5 * (c == 1) + of ** 2
    return float(a), float(b)


def branin(cfg):
    x1, x2 = cfg["x1"], cfg["x2"]
    a, b, c, r, s, t = 1.0, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6.0, 10.0, 1 / (8 * np.pi)
    return float(a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s)


def objective_discrete(cfg):
    return float((cfg["a"] - 11) ** 2 / 10.0 + abs(np.log2(cfg["b"]) - 3) * 2 + [3, 0, 1, 2][cfg["c"]] + cfg["d"])


def to_native(v):
    return v.item() if hasattr(v, "item") else v


def sample_data_array(space, n, fn, objectives):
    cfgs = space.get_random_configuration(size=n, use_priors=False)
    data = {k: [] for k in space.get_input_parameters()}
    for o in objectives:
        data[o] = []
    for c in cfgs:
        for k in space.get_input_parameters():
            data[k].append(to_native(c[k]))
        vals = fn(c)
        if not isinstance(vals, tuple):
            vals = (vals,)
        for o, v in zip(objectives, vals):
            data[o].append(v)
    return data


def fit(config, space, data, objectives):
    with ref_harness.logger_stdout():
        models, _, _ = ns.models.generate_mono_output_regression_models(
            data, space, space.get_input_parameters(), objectives, 1.0, config, model_type="random_forest", number_of_cpus=1)
    return models


def rf_case(config, fn, n_train, n_cand, seed, fname):
    from oracle import oracle
    config = validated(config)
    np.random.seed(seed)
    random.seed(seed)
    space = ns.space.Space(config)
    objectives = config["optimization_objectives"]
    inputs = space.get_input_parameters()
    data = sample_data_array(space, n_train, fn, objectives)
    models = fit(config, space, data, objectives)

    cand = space.get_random_configuration(size=n_cand, use_priors=False, return_as_array=True)
    configurations = ns.utility_functions.array_to_list_of_dicts(cand, inputs)
    bufferx = [tuple(to_native(c[k]) for k in inputs) for c in configurations]

    out = {}
    out["params_json"] = np.array(json.dumps(params_desc(config)))
    out["objectives_json"] = np.array(json.dumps(objectives))
    out["config_json"] = np.array(json.dumps(config))
    out["candidates"] = np.array([[float(v) for v in row] for row in bufferx], dtype=np.float64)
    out["data_inputs"] = np.array([[float(data[k][r]) for k in inputs] for r in range(n_train)], dtype=np.float64)
    for o in objectives:
        out["data_" + o] = np.array(data[o], dtype=np.float64)

    with ref_harness.logger_stdout():
        enc = ns.models.preprocess_data_buffer(bufferx, space)
        out["encoded"] = np.array(enc, dtype=np.float64)
        for o in objectives:
            fa = oracle.forest_arrays_from_model(models[o])
            out.update(fa.to_npz_dict("forest_%s_" % o))
            out["leaves_" + o] = models[o].get_leaves_per_sample(enc, space)
            out["sk_predict_" + o] = models[o].predict(enc)
        mean, var = ns.models.compute_model_mean_and_uncertainty(bufferx, models, "random_forest", space, var=True)
        _, std = ns.models.compute_model_mean_and_uncertainty(bufferx, models, "random_forest", space, var=False)
        for o in objectives:
            out["mean_" + o] = mean[o]
            out["var_" + o] = var[o]
            out["std_" + o] = std[o]

        weights = {o: w for o, w in zip(objectives, [0.37, 0.63][: len(objectives)] if len(objectives) > 1 else [1.0])}
        limits = {o: [float("inf"), float("-inf")] for o in objectives}
        for o in objectives:
            limits[o] = [min(data[o]), max(data[o])]
        # make the running-limit update matter: shrink the stored limits a little
        stored_limits = {o: [limits[o][0] + 0.1 * (limits[o][1] - limits[o][0]), limits[o][1]] for o in objectives}
        out["weights"] = np.array([weights[o] for o in objectives])
        out["stored_limits"] = np.array([stored_limits[o] for o in objectives])
        iteration_number = 7
        out["iteration_number"] = np.int64(iteration_number)
        for scal in ("linear", "tchebyshev", "modified_tchebyshev"):
            s, lim = ns.utility_functions.compute_data_array_scalarization(data, weights, copy.deepcopy(stored_limits), scal)
            out["data_scalarization_" + scal] = np.array(s)
            out["data_scalarization_limits_" + scal] = np.array([lim[o] for o in objectives])
            for acq in ("EI", "UCB", "TS"):
                vals, feas = ns.random_scalarizations.run_acquisition_function(
                    acq, configurations, weights, models, space, scal, copy.deepcopy(stored_limits), iteration_number,
                    data, "random_forest", None, 1)
                out["acq_%s_%s" % (acq, scal)] = np.array(vals, dtype=np.float64)
                assert feas == [1] * len(vals)
        # get_min_configurations in index form
        vals = out["acq_EI_tchebyshev"]
        d = {"__index": list(range(len(vals))), "scal": list(vals)}
        for k in (1, 20, 64):
            best = ns.utility_functions.get_min_configurations(d, k, "scal")
            out["topk_idx_%d" % k] = np.array(best["__index"], dtype=np.int64)
            out["topk_val_%d" % k] = np.array(best["scal"], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, fname), **out)
    print(fname, {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:6]},
          "distinct EI", len(np.unique(out["acq_EI_tchebyshev"])))


def pareto_case():
    out = {}
    f = ns.compute_pareto.sequential_is_pareto_efficient
    rng = np.random.default_rng(7)
    cases = {
        "uniform3": rng.random((4000, 3)),
        "uniform2": rng.random((4000, 2)),
        "uniform4": rng.random((1500, 4)),
        "ties3": np.floor(rng.random((3000, 3)) * 16) / 16,
        "ties2": np.floor(rng.random((3000, 2)) * 32) / 32,
        "anticorr2": None,
        "dups3": None,
        "nan3": None,
        "single": np.array([[1.0, 2.0, 3.0]]),
        "quirk3": np.array([[1.0, 5.0, 9.0], [1.0, 4.0, 10.0]]),
        "quirk3_swapped": np.array([[1.0, 4.0, 10.0], [1.0, 5.0, 9.0]]),
    }
    x = rng.random(2000)
    cases["anticorr2"] = np.stack([x, 1 - x + 0.01 * rng.random(2000)], 1)
    base = rng.random((500, 3))
    cases["dups3"] = np.concatenate([base, base[:200], base[100:300]], 0)
    n3 = rng.random((1500, 3))
    n3[rng.integers(0, 1500, 25), rng.integers(0, 3, 25)] = np.nan
    n3[7] = np.nan
    cases["nan3"] = n3
    for k, c in cases.items():
        with np.errstate(all="ignore"):
            out["costs_" + k] = c
            out["mask_" + k] = f(c.copy())
    # the reference's own known-answer cases (tests/test_system.py + tests/data/*.csv)
    import csv
    rows = list(csv.DictReader(open(os.path.join(REF, "tests/data/BlackScholes_exhaustive_search_data.csv"))))
    valid = [r for r in rows if r["Valid"].strip().lower() == "true"]
    costs = np.array([[float(r["ALMs"]), float(r["Cycles"])] for r in valid])
    out["costs_blackscholes"] = costs
    out["mask_blackscholes"] = f(costs.copy())
    gold = list(csv.DictReader(open(os.path.join(REF, "tests/data/BlackScholes_real_pareto.csv"))))
    out["kat_blackscholes"] = np.array(sorted([[float(r["ALMs"]), float(r["Cycles"])] for r in gold]))
    got = np.array(sorted(costs[out["mask_blackscholes"]].tolist()))
    assert got.shape == out["kat_blackscholes"].shape and np.array_equal(got, out["kat_blackscholes"]), "BlackScholes KAT"

    # 100x100 Branin grid of the quick-start example -> branin_grid_search_pareto.csv
    sys.path.insert(0, os.path.join(REF, "example_scenarios", "quick_start"))
    gold = list(csv.DictReader(open(os.path.join(REF, "tests/data/branin_grid_search_pareto.csv"))))
    keys = list(gold[0].keys())
    out["kat_branin_grid_cols"] = np.array(json.dumps(keys))
    out["kat_branin_grid"] = np.array([[float(r[k]) for k in keys] for r in gold])
    gold = list(csv.DictReader(open(os.path.join(REF, "tests/data/ordinal_branin_real_pareto.csv"))))
    keys = list(gold[0].keys())
    out["kat_ordinal_branin_cols"] = np.array(json.dumps(keys))
    out["kat_ordinal_branin"] = np.array([[float(r[k]) for k in keys] for r in gold])
    np.savez_compressed(os.path.join(HERE, "pareto.npz"), **out)
    print("pareto.npz", {k: int(v.sum()) for k, v in out.items() if k.startswith("mask_")})


def local_search_case():
    """local_search() on a discrete space (deterministic given the seeds): frozen models + picked configuration."""
    from oracle import oracle
    config = validated(DISCRETE)
    out = {}
    np.random.seed(11)
    random.seed(11)
    space = ns.space.Space(config)
    objectives = config["optimization_objectives"]
    inputs = space.get_input_parameters()
    data = sample_data_array(space, 60, objective_discrete, objectives)
    models = fit(config, space, data, objectives)
    fa = oracle.forest_arrays_from_model(models["Value"])
    out.update(fa.to_npz_dict("forest_Value_"))
    out["params_json"] = np.array(json.dumps(params_desc(config)))
    out["config_json"] = np.array(json.dumps(config))
    out["data_inputs"] = np.array([[float(data[k][r]) for k in inputs] for r in range(60)])
    out["data_Value"] = np.array(data["Value"])
    fast = {}
    for r in range(60):
        fast[space.get_unique_hash_string_from_values({k: data[k][r] for k in inputs})] = r
    weights = {"Value": 1.0}
    limits = {"Value": [min(data["Value"]), max(data["Value"])]}
    for trial, seed in enumerate((3, 4, 5)):
        for acq in ("EI", "UCB"):
            np.random.seed(seed)
            random.seed(seed)
            state = np.random.get_state()
            d2 = copy.deepcopy(data)
            scal, _ = ns.utility_functions.compute_data_array_scalarization(d2, weights, copy.deepcopy(limits), "tchebyshev")
            d2[config["scalarization_key"]] = scal.tolist()
            cfg2 = copy.deepcopy(config)
            cfg2["acquisition_function"] = acq
            with ref_harness.logger_stdout():
                best = ns.random_scalarizations.random_scalarizations(
                    cfg2, d2, space, fast, models, 3 + trial, weights, copy.deepcopy(limits), None, None, "local_search")
            out["best_%s_%d" % (acq, seed)] = np.array([float(best[k]) for k in inputs])
            # the pools the reference drew (same RNG state -> same draws), for pool-level parity
            np.random.set_state(state)
            uni = space.get_random_configuration(size=config["local_search_random_points"], use_priors=False, return_as_array=True)
            pri = space.get_random_configuration(size=config["local_search_random_points"], use_priors=True, return_as_array=True)
            out["pool_uniform_%s_%d" % (acq, seed)] = uni.astype(np.float64)
            out["pool_prior_%s_%d" % (acq, seed)] = pri.astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "local_search.npz"), **out)
    print("local_search.npz", {k: v.tolist() for k, v in out.items() if k.startswith("best_")})



PRIORS = {
    "application_name": "golden_priors",
    "optimization_objectives": ["obj_a"],
    "optimization_iterations": 20,
    "models": {"model": "random_forest", "number_of_trees": 8},
    "custom_gaussian_prior_means": [0.3],
    "custom_gaussian_prior_stds": [0.15],
    "input_parameters": {
        "r_gauss": {"parameter_type": "real", "values": [-5, 10], "parameter_default": 0, "prior": "gaussian"},
        "r_decay": {"parameter_type": "real", "values": [0, 1], "parameter_default": 0.5, "prior": "decay"},
        "r_list": {"parameter_type": "real", "values": [2, 6], "parameter_default": 3,
                   "prior": [1, 2, 4, 8, 4, 2, 1, 1, 3, 1]},
        "r_cg": {"parameter_type": "real", "values": [0, 1], "parameter_default": 0.5, "prior": "custom_gaussian"},
        "i_exp": {"parameter_type": "integer", "values": [0, 20], "parameter_default": 3, "prior": "exponential"},
        "i_list": {"parameter_type": "integer", "values": [1, 4], "parameter_default": 2, "prior": [0.1, 0.2, 0.3, 0.4]},
        "o_gauss": {"parameter_type": "ordinal", "values": [1, 2, 4, 8, 16, 32, 64, 128], "parameter_default": 4,
                    "prior": "gaussian"},
        "o_list": {"parameter_type": "ordinal", "values": [0.5, 1.5, 2.5], "parameter_default": 1.5,
                   "prior": [0.5, 0.25, 0.25]},
        "c_list": {"parameter_type": "categorical", "values": ["a", "b", "c"], "parameter_default": "a",
                   "prior": [0.2, 0.5, 0.3]},
        "c_uni": {"parameter_type": "categorical", "values": ["p", "q"], "parameter_default": "q"},
    },
}


def objective_priors(cfg):
    return float((cfg["r_gauss"] - 2.0) ** 2 + 3 * cfg["r_decay"] + cfg["r_list"] + 5 * (cfg["r_cg"] - 0.3) ** 2 +
                 0.3 * cfg["i_exp"] + cfg["i_list"] + np.log2(cfg["o_gauss"]) * (1 + cfg["c_list"]) +
                 2 * cfg["o_list"] - 3 * cfg["c_uni"])


def priors_case():
    """Every prior form of space.py through the reference's prior-guided acquisition functions."""
    from oracle import oracle
    config = validated(PRIORS)
    np.random.seed(31)
    random.seed(31)
    space = ns.space.Space(config)
    objectives = config["optimization_objectives"]
    inputs = space.get_input_parameters()
    data = sample_data_array(space, 80, objective_priors, objectives)
    models = fit(config, space, data, objectives)
    uni = space.get_random_configuration(size=300, use_priors=False, return_as_array=True)
    pri = space.get_random_configuration(size=300, use_priors=True, return_as_array=True)
    cand = np.concatenate([uni, pri], axis=0)
    # boundary values: pdf = inf for the decay / exponential densities (capped at 1242 for reals only)
    cand[0, inputs.index("r_decay")] = 0.0
    cand[1, inputs.index("i_exp")] = 20
    cand[2, inputs.index("r_gauss")] = -5.0
    cand[3, inputs.index("r_list")] = 6.0
    configurations = [{k: to_native(c[k]) for k in inputs} for c in ns.utility_functions.array_to_list_of_dicts(cand, inputs)]
    out = {}
    out["config_json"] = np.array(json.dumps(config))
    out["candidates"] = np.array([[float(c[k]) for k in inputs] for c in configurations], dtype=np.float64)
    out["data_inputs"] = np.array([[float(data[k][r]) for k in inputs] for r in range(80)], dtype=np.float64)
    out["data_obj_a"] = np.array(data["obj_a"], dtype=np.float64)
    out.update(oracle.forest_arrays_from_model(models["obj_a"]).to_npz_dict("forest_obj_a_"))
    with ref_harness.logger_stdout():
        out["prior_w1"] = np.array(ns.prior_optimization.compute_probability_from_prior(
            configurations, space, {"obj_a": 1.0}), dtype=np.float64)
        out["prior_w07"] = np.array(ns.prior_optimization.compute_probability_from_prior(
            configurations, space, {"obj_a": 0.7}), dtype=np.float64)
        weights = {"obj_a": 1.0}
        limits = {"obj_a": [min(data["obj_a"]), max(data["obj_a"])]}
        for acq in ("EI", "UCB", "TS"):
            for it in (1, 7):
                vals, feas = ns.pibo.prior_weighted_acquisition_function(
                    copy.deepcopy(configurations), space, ns.random_scalarizations.run_acquisition_function, acq,
                    weights, copy.deepcopy(limits), it, 2.0, models, None, data, "random_forest", "linear")
                out["pibo_%s_it%d" % (acq, it)] = np.array(vals, dtype=np.float64)
    # BOPrO: compute_EI_from_posteriors calls get_min()/get_max() on every parameter (prior_optimization.py:189-197),
    # which CategoricalParameter does not have -> the reference only runs it on spaces without categoricals
    cfg_nc = copy.deepcopy(PRIORS)
    for k in ("c_list", "c_uni"):
        del cfg_nc["input_parameters"][k]
    # Beta(1, 0.8) has an infinite density at the upper end, which an integer parameter reaches: the prior limits
    # become [.., inf] and every normalised value NaN.  Use the bell-shaped density here (zero at both ends).
    cfg_nc["input_parameters"]["i_exp"]["prior"] = "gaussian"
    cfg_nc = validated(cfg_nc)
    np.random.seed(32)
    random.seed(32)
    space_nc = ns.space.Space(cfg_nc)
    inputs_nc = space_nc.get_input_parameters()
    fn_nc = lambda c: objective_priors(dict(c, c_list=1, c_uni=0))
    data_nc = sample_data_array(space_nc, 80, fn_nc, objectives)
    models_nc = fit(cfg_nc, space_nc, data_nc, objectives)
    cand_nc = np.concatenate([space_nc.get_random_configuration(size=300, use_priors=False, return_as_array=True),
                              space_nc.get_random_configuration(size=300, use_priors=True, return_as_array=True)], axis=0)
    conf_nc = [{k: to_native(c[k]) for k in inputs_nc}
               for c in ns.utility_functions.array_to_list_of_dicts(cand_nc, inputs_nc)]
    out["nc_config_json"] = np.array(json.dumps(cfg_nc))
    out["nc_candidates"] = np.array([[float(c[k]) for k in inputs_nc] for c in conf_nc], dtype=np.float64)
    out["nc_data_obj_a"] = np.array(data_nc["obj_a"], dtype=np.float64)
    out.update(oracle.forest_arrays_from_model(models_nc["obj_a"]).to_npz_dict("nc_forest_obj_a_"))
    with ref_harness.logger_stdout():
        weights = {"obj_a": 1.0}
        limits = {"obj_a": [min(data_nc["obj_a"]), max(data_nc["obj_a"])]}
        thr = {"obj_a": float(np.quantile(data_nc["obj_a"], 0.05))}
        out["bopro_threshold"] = np.array([thr["obj_a"]])
        # (a) raw prior, no normalisation limits
        vals, feas = ns.prior_optimization.compute_EI_from_posteriors(
            copy.deepcopy(conf_nc), space_nc, weights, copy.deepcopy(limits), thr, 5, 10, models_nc, None,
            "random_forest", None)
        out["bopro_plain"] = np.array(vals, dtype=np.float64)
        # (b) prior normalised with running limits, posterior normalised too (the constrained-run code path)
        gl, pl = [0.5, 0.6], [float("inf"), float("-inf")]
        vals, feas = ns.prior_optimization.compute_EI_from_posteriors(
            copy.deepcopy(conf_nc), space_nc, weights, copy.deepcopy(limits), thr, 5, 10, models_nc, None,
            "random_forest", gl, posterior_normalization_limits=pl)
        out["bopro_norm"] = np.array(vals, dtype=np.float64)
        out["bopro_norm_prior_limits"] = np.array(gl, dtype=np.float64)
        out["bopro_norm_posterior_limits"] = np.array(pl, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "priors.npz"), **out)
    print("priors.npz", {k: (v.shape, float(np.nanmin(v)), float(np.nanmax(v))) for k, v in out.items()
                         if k.startswith(("prior_", "pibo_", "bopro_"))})

# ---------------------------------------------------------------------------------------------------------------------
# local_search() trajectories on spaces WITH real / integer parameters (C1 = the bundled Branin scenario shape, and the
# mixed space) and on the discrete space with a different training set per seed.  The neighbourhoods of real / integer
# parameters draw truncated normals from numpy's global stream (local_search.py:138-157), so the RNG state right after
# the two pools were drawn is frozen next to the pools: a replay that restores it consumes the very same numbers.
# ---------------------------------------------------------------------------------------------------------------------
def _traj_one(out, tag, config, space, models, data, objectives, weights, limits, seed, acq, scal, iteration,
              n_points, n_starts):
    inputs = space.get_input_parameters()
    n_train = len(data[objectives[0]])
    fast = {}
    for r in range(n_train):
        fast[space.get_unique_hash_string_from_values({k: data[k][r] for k in inputs})] = r
    d2 = copy.deepcopy(data)
    s, _ = ns.utility_functions.compute_data_array_scalarization(d2, weights, copy.deepcopy(limits), scal)
    key = config["scalarization_key"]
    d2[key] = s.tolist()
    params = {"regression_models": models, "iteration_number": iteration, "data_array": d2,
              "classification_model": None, "param_space": space, "objective_weights": weights,
              "model_type": "random_forest", "objective_limits": copy.deepcopy(limits), "acquisition_function": acq,
              "scalarization_method": scal, "number_of_cpus": 1}
    pools, states, colorder = [], [], []
    orig_sample = space.get_random_configuration
    orig_numpy = ns.local_search.dict_of_lists_to_numpy

    def sample(*a, **kw):
        r = orig_sample(*a, **kw)
        if kw.get("return_as_array"):
            pools.append(np.array(r, dtype=np.float64))
            states.append(np.random.get_state())
        return r

    def to_numpy(d, return_col_of_key=False):
        r = orig_numpy(d, return_col_of_key=return_col_of_key)
        if return_col_of_key:
            colorder.append([k for k, _ in sorted(r[1].items(), key=lambda kv: kv[1])])
        return r

    space.get_random_configuration = sample
    ns.local_search.dict_of_lists_to_numpy = to_numpy
    np.random.seed(seed)
    random.seed(seed)
    try:
        with ref_harness.logger_stdout():
            data_array, best = ns.local_search.local_search(
                n_starts, n_points, space, fast, False, ns.random_scalarizations.run_acquisition_function, params, key, 1,
                previous_points=d2, profiling=None)
    finally:
        space.get_random_configuration = orig_sample
        ns.local_search.dict_of_lists_to_numpy = orig_numpy
    assert len(pools) == 2 and len(colorder) == 1
    pre = "%s_" % tag
    out[pre + "pool_uniform"] = pools[0]
    out[pre + "pool_prior"] = pools[1]
    st = states[1]
    assert st[0] == "MT19937"
    out[pre + "rng_keys"] = np.asarray(st[1], dtype=np.uint32)
    out[pre + "rng_pos"] = np.array([st[2], st[3]], dtype=np.int64)
    out[pre + "rng_gauss"] = np.array([st[4]], dtype=np.float64)
    out[pre + "dedup_columns"] = np.array(json.dumps(colorder[0]))
    out[pre + "best"] = np.array([float(best[k]) for k in inputs])
    out[pre + "data_array"] = np.array([[float(v) for v in data_array[k]] for k in inputs + [key]], dtype=np.float64)
    out[pre + "meta"] = np.array(json.dumps({"seed": seed, "acq": acq, "scal": scal, "iteration": iteration,
                                             "n_points": n_points, "n_starts": n_starts}))
    return best


def local_search_traj_case():
    out = {}
    summary = {}
    plans = [
        # tag, scenario, objective, n_train, pool size, starts, [(seed, acq, scal)]
        ("branin", dict(BRANIN, models={"model": "random_forest", "number_of_trees": 10}), branin, 25, 1000, 10,
         [(0, "EI", "tchebyshev"), (1, "EI", "tchebyshev"), (2, "UCB", "linear")]),
        ("mixed", MIXED, objective_mixed, 70, 400, 5,
         [(0, "EI", "tchebyshev"), (1, "UCB", "modified_tchebyshev"), (2, "TS", "linear")]),
        ("discrete", DISCRETE, objective_discrete, 50, 300, 4,
         [(0, "EI", "tchebyshev"), (1, "UCB", "linear"), (2, "EI", "linear"), (3, "TS", "tchebyshev")]),
    ]
    for name, scenario, fn, n_train, n_points, n_starts, runs in plans:
        config = validated(scenario)
        objectives = config["optimization_objectives"]
        out[name + "_config_json"] = np.array(json.dumps(config))
        out[name + "_params_json"] = np.array(json.dumps(params_desc(config)))
        for seed, acq, scal in runs:
            tag = "%s_s%d" % (name, seed)
            np.random.seed(1000 + seed)            # a different training set (and forest) per run
            random.seed(1000 + seed)
            space = ns.space.Space(config)
            inputs = space.get_input_parameters()
            data = sample_data_array(space, n_train, fn, objectives)
            models = fit(config, space, data, objectives)
            for o in objectives:
                out.update(oracle_forest(models[o]).to_npz_dict("%s_forest_%s_" % (tag, o)))
                out["%s_data_%s" % (tag, o)] = np.array(data[o], dtype=np.float64)
            out[tag + "_data_inputs"] = np.array([[float(data[k][r]) for k in inputs] for r in range(n_train)])
            weights = {o: w for o, w in zip(objectives, [0.37, 0.63] if len(objectives) > 1 else [1.0])}
            limits = {o: [min(data[o]), max(data[o])] for o in objectives}
            out[tag + "_weights"] = np.array([weights[o] for o in objectives])
            best = _traj_one(out, tag, config, space, models, data, objectives, weights, limits, seed, acq, scal,
                             3 + seed, n_points, n_starts)
            summary[tag] = [float(best[k]) for k in inputs]
    np.savez_compressed(os.path.join(HERE, "local_search_traj.npz"), **out)
    print("local_search_traj.npz", summary)


def oracle_forest(model):
    from oracle import oracle
    return oracle.forest_arrays_from_model(model)


# ---------------------------------------------------------------------------------------------------------------------
# GP: the reference delegates to GPy (models.py:441-459, 997-1024), which is not installable here.  The restatement in
# oracle/oracle.py is pinned to an INDEPENDENT exact-inference implementation instead: scikit-learn's
# GaussianProcessRegressor with ConstantKernel * Matern(nu=2.5, ARD) [+ WhiteKernel], fixed hyper-parameters
# (optimizer=None), normalize_y=True (population std, like GPy's Standardize), alpha = 1e-8 (the jitter GPy's
# exact_gaussian_inference adds to the noise variance).  sklearn evaluates the kernel from explicit pairwise distances
# (cdist), not GPy's GEMM form, and solves with its own Cholesky, so agreement is a genuine cross-check.
# ---------------------------------------------------------------------------------------------------------------------
def gp_case():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel, RBF
    sys.path.insert(0, ROOT)
    from bench_support import workloads
    out = {}
    cases = [
        ("c2_noise", 256, 6, 1e-4, 2, "matern52", 1.0, 1500),
        ("c2_nonoise", 256, 6, 0.0, 2, "matern52", 1.0, 1500),
        ("c5gp", 512, 32, 1e-4, 12, "matern52", 3.0, 1000),
        ("small_rbf", 60, 4, 1e-3, 21, "rbf", 1.0, 500),
        ("sf2", 200, 5, 1e-2, 22, "matern52", 1.0, 500),
    ]
    for tag, n, D, noise, seed, kernel, scale, N in cases:
        rng = np.random.default_rng(seed)
        X = rng.random((n, D))
        if D == 6:
            y = workloads.hartmann6(X)
        elif D == 32:
            y = np.sin(X[:, :8].sum(1)) + (X[:, 8:] ** 2).sum(1) * 0.1
        else:
            y = np.sin(3 * X[:, 0]) + (X ** 2).sum(1)
        ls = np.random.default_rng(seed + 1).uniform(0.3, 1.0, D) * scale
        sf2 = 1.0 if tag != "sf2" else 2.7
        base = Matern(length_scale=ls, nu=2.5) if kernel == "matern52" else RBF(length_scale=ls)
        k = ConstantKernel(sf2, constant_value_bounds="fixed") * base
        if noise > 0:
            k = k + WhiteKernel(noise, noise_level_bounds="fixed")
        gpr = GaussianProcessRegressor(kernel=k, alpha=1e-8, optimizer=None, normalize_y=True)
        gpr.fit(X, y)
        Xs = np.random.default_rng(seed + 2).random((N, D))
        Xs[:5] = X[:5]                     # candidates ON training points: the ill-conditioned end of the variance
        Xs[5:10] = X[5:10] + 1e-4
        mean, std = gpr.predict(Xs, return_std=True)
        out[tag + "_X"] = X
        out[tag + "_y"] = y
        out[tag + "_lengthscale"] = ls
        out[tag + "_hyper"] = np.array([sf2, noise])
        out[tag + "_kernel"] = np.array(kernel)
        out[tag + "_candidates"] = Xs
        out[tag + "_sk_mean"] = mean
        out[tag + "_sk_var"] = std ** 2
    np.savez_compressed(os.path.join(HERE, "gp_sklearn.npz"), **out)
    print("gp_sklearn.npz", {k: v.shape for k, v in out.items() if k.endswith("_sk_mean")})



def neighbors_case():
    """get_neighbors (local_search.py:83-161) on the discrete and the mixed scenario, RNG seeded per call."""
    out = {}
    for tag, cfg in (("discrete", DISCRETE), ("mixed", MIXED)):
        config = validated(cfg)
        space = ns.space.Space(config)
        inputs = space.get_input_parameters()
        np.random.seed(21)
        starts = space.get_random_configuration(size=6, use_priors=False, return_as_array=True)
        out["params_json_" + tag] = np.array(json.dumps(params_desc(config)))
        out["config_json_" + tag] = np.array(json.dumps(config))
        out["starts_" + tag] = starts.astype(np.float64)
        for r in range(len(starts)):
            conf = {k: to_native(starts[r, j]) for j, k in enumerate(inputs)}
            np.random.seed(100 + r)
            nb = ns.local_search.get_neighbors(conf, space)
            out["neighbors_%s_%d" % (tag, r)] = np.array([[float(v) for v in row] for row in nb], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "neighbors.npz"), **out)
    print("neighbors.npz", {k: v.shape for k, v in out.items() if k.startswith("neighbors_")})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "neighbors":
        neighbors_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "priors":
        priors_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "traj":
        local_search_traj_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "gp":
        gp_case()
        sys.exit(0)
    rf_case(MIXED, objective_mixed, 90, 700, 1, "rf_mixed.npz")
    rf_case(BRANIN, branin, 40, 1500, 2, "rf_branin.npz")
    pareto_case()
    local_search_case()
    neighbors_case()
    priors_case()
    local_search_traj_case()
    gp_case()
