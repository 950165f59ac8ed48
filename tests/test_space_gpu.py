"""GPU parity of the search-space kernels (csrc/hm_space.cu) through the reference-facing API:
K5 hm_encode (models.preprocess_data_array), K7 hm_prior_prob (compute_probability_from_prior), the piBO weighting
(pibo.prior_weighted_acquisition_function), BOPrO (prior_optimization.compute_EI_from_posteriors) against golden
vectors produced by the unmodified reference (tests/golden/make_golden.py), and K6 hm_sample_pool against the analytic
distributions of space.py's randomly_select / randomly_select_uniform (the reference draws from numpy's global
Mersenne-Twister stream, which a counter-based device generator cannot and need not reproduce)."""
import copy
import json
import os

import numpy as np
import pytest
from scipy import stats

from conftest import GOLDEN
from test_api_gpu import FrozenForest

pytestmark = pytest.mark.gpu

REL = 1e-5


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def space_of(z, key="config_json"):
    from bench_support.space_lite import SpaceLite
    return SpaceLite(json.loads(str(z[key])))


def close(got, want, rel=REL):
    """allclose with infinities / NaNs required to coincide exactly"""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    fin = np.isfinite(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)])
    assert np.allclose(got[fin], want[fin], rtol=rel, atol=0), np.max(np.abs(got[fin] - want[fin]) / np.abs(want[fin]).clip(1e-300))
    return True


# ---- K5 ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
def test_device_encode_matches_reference_preprocess(name):
    from hypermapper_b200.space_device import DeviceSpace
    z = load(name)
    ds = DeviceSpace(space_of(z))
    x32, x64 = ds.encode(z["candidates"], want32=True, want64=True)
    assert np.array_equal(x64.cpu().numpy().T, z["encoded"])                                   # bit-exact
    assert np.array_equal(x32.cpu().numpy().T, z["encoded"].astype(np.float32))                # sklearn's cast
    assert ds.n_encoded == z["encoded"].shape[1]


def test_device_encode_api_forms_and_errors():
    from hypermapper_b200 import models
    z = load("rf_mixed.npz")
    space = space_of(z)
    inputs = space.get_input_parameters()
    cand = z["candidates"][:50]
    as_tuples = [tuple(r) for r in cand]
    as_dicts = [dict(zip(inputs, r)) for r in cand]
    want = z["encoded"][:50]
    for form in (cand, as_tuples, as_dicts):
        assert np.array_equal(models.encode_to_device(form, space, "float64").cpu().numpy().T, want)
    assert models.preprocess_data_buffer(as_tuples, space) == [tuple(r) for r in want]
    bad = cand.copy()
    bad[1, 2] = 3.0                  # not one of the ordinal's values: list.index raises in the reference
    with pytest.raises(ValueError):
        models.encode_to_device(bad, space, "float32")
    bad = cand.copy()
    bad[4, 3] = 7.0                  # unknown category: OneHotEncoder raises in the reference
    with pytest.raises(ValueError):
        models.encode_to_device(bad, space, "float32")
    assert models.encode_to_device(cand[:0], space, "float32").shape == (want.shape[1], 0)
    # an integer parameter with 2^31 or more values: the device generator draws value indices with 32-bit arithmetic, so
    # the space is refused when its handle is built instead of sampling garbage
    from hypermapper_b200 import _lib
    from hypermapper_b200.space_device import DeviceSpace
    from bench_support.space_lite import SpaceLite
    huge = SpaceLite({"optimization_objectives": ["v"], "input_parameters": {
        "i": {"parameter_type": "integer", "values": [0, 2 ** 31]}, "x": {"parameter_type": "real", "values": [0, 1]}}})
    with pytest.raises(_lib.HmError):
        DeviceSpace(huge)


def test_device_encode_zscore_matches_host_encoder():
    """normalize_inputs: (x - mean) / std in fp64 (models.py:966-971); the statistics are read at encode time."""
    from hypermapper_b200.encoding import Encoder
    from hypermapper_b200.space_device import device_space_for
    z = load("rf_mixed.npz")
    cfg = json.loads(str(z["config_json"]))
    cfg["normalize_inputs"] = True
    from bench_support.space_lite import SpaceLite
    space = SpaceLite(cfg)
    for rep, (m, s) in enumerate(((1.25, 3.5), (-0.3, 0.07))):
        for p in ("x_real", "x_int"):
            space.set_parameter_mean(p, m + rep)
            space.set_parameter_std(p, s)
        want = Encoder(space).encode(z["candidates"])
        _, x64 = device_space_for(space).encode(z["candidates"], want32=False, want64=True)
        assert np.array_equal(x64.cpu().numpy(), want)


# ---- K7 / piBO / BOPrO ---------------------------------------------------------------------------------------------
def test_prior_probability_matches_reference():
    from hypermapper_b200 import prior_optimization as po
    z = load("priors.npz")
    space = space_of(z)
    inputs = space.get_input_parameters()
    confs = [dict(zip(inputs, r)) for r in z["candidates"]]
    for key, w in (("prior_w1", 1.0), ("prior_w07", 0.7)):
        got = po.compute_probability_from_prior(confs, space, {"obj_a": w})
        assert isinstance(got, list) and len(got) == len(confs)
        close(got, z[key], rel=1e-9)
    assert np.isinf(z["prior_w1"]).any() and (z["prior_w1"] == 0).any()       # the fixture covers the edge values


def _priors_models(z, prefix=""):
    return {"obj_a": FrozenForest(z, prefix + "forest_obj_a_")}


@pytest.mark.parametrize("acq", ["EI", "UCB", "TS"])
@pytest.mark.parametrize("it", [1, 7])
def test_pibo_prior_weighted_acquisition_matches_reference(acq, it):
    from hypermapper_b200 import pibo, random_scalarizations as rs
    z = load("priors.npz")
    space = space_of(z)
    inputs = space.get_input_parameters()
    confs = [dict(zip(inputs, r)) for r in z["candidates"]]
    data = {p: z["data_inputs"][:, j].tolist() for j, p in enumerate(inputs)}
    data["obj_a"] = z["data_obj_a"].tolist()
    limits = {"obj_a": [min(data["obj_a"]), max(data["obj_a"])]}
    vals, feas = pibo.prior_weighted_acquisition_function(
        confs, space, rs.run_acquisition_function, acq, {"obj_a": 1.0}, copy.deepcopy(limits), it, 2.0,
        _priors_models(z), None, data, "random_forest", "linear")
    assert feas == [1] * len(confs)
    close(vals, z["pibo_%s_it%d" % (acq, it)])


def test_bopro_posterior_scores_match_reference():
    from hypermapper_b200 import prior_optimization as po
    z = load("priors.npz")
    space = space_of(z, "nc_config_json")
    inputs = space.get_input_parameters()
    confs = [dict(zip(inputs, r)) for r in z["nc_candidates"]]
    models = _priors_models(z, "nc_")
    thr = {"obj_a": float(z["bopro_threshold"][0])}
    lim = {"obj_a": [float(z["nc_data_obj_a"].min()), float(z["nc_data_obj_a"].max())]}
    vals, feas = po.compute_EI_from_posteriors(copy.deepcopy(confs), space, {"obj_a": 1.0}, lim, thr, 5, 10, models,
                                               None, "random_forest", None)
    assert list(feas) == [1] * len(confs)
    close(vals, z["bopro_plain"])
    gl, pl = [0.5, 0.6], [float("inf"), float("-inf")]
    vals, _ = po.compute_EI_from_posteriors(copy.deepcopy(confs), space, {"obj_a": 1.0}, lim, thr, 5, 10, models,
                                            None, "random_forest", gl, posterior_normalization_limits=pl)
    close(vals, z["bopro_norm"])
    close(gl, z["bopro_norm_prior_limits"], rel=1e-9)          # the running limits are updated in place (:173-178)
    close(pl, z["bopro_norm_posterior_limits"])


def test_local_search_runs_with_prior_weighted_plugins():
    """The two prior-guided plug-ins drive local_search on the device path and return an unseen configuration."""
    from hypermapper_b200 import pibo, prior_optimization as po
    from hypermapper_b200.utility_functions import compute_data_array_scalarization
    z = load("priors.npz")
    for key, prefix, runner in (("config_json", "", pibo.run_pibo), ("nc_config_json", "nc_", po.prior_guided_optimization)):
        cfg = json.loads(str(z[key]))
        cfg.update({"local_search_random_points": 2000, "local_search_starting_points": 3, "number_of_cpus": 1,
                    "scalarization_key": "scalarization", "scalarization_method": "linear",
                    "acquisition_function": "EI", "prior_beta": -1, "prior_floor": 1e-6,
                    "model_posterior_weight": 10, "posterior_computation_lower_limit": 1e-8,
                    "model_good_quantile": 0.05, "prior_limit_estimation_points": 200})
        cfg.setdefault("models", {})["model"] = "random_forest"
        # get_neighbors maps 0 -> round(0 * size - 0.5 + min) (space.py:437-453): with min = 1 that is 0, outside the
        # range, and the reference's list.index raises for a "distribution" prior.  [0, 4] has no such boundary case.
        cfg["input_parameters"]["i_list"].update({"values": [0, 4], "prior": [0.1, 0.2, 0.3, 0.2, 0.2]})
        from bench_support.space_lite import SpaceLite
        space = SpaceLite(cfg)
        inputs = space.get_input_parameters()
        np.random.seed(5)
        rows = z[prefix + "candidates"][:40]
        data = {p: [space.get_input_parameters_objects()[p].get_values()[int(np.searchsorted(
            np.asarray(space.get_input_parameters_objects()[p].get_values(), dtype=float), v))]
            if space.get_type(p) == "ordinal" else (float(v) if space.get_type(p) == "real" else int(v))
            for v in rows[:, j]] for j, p in enumerate(inputs)}
        data["obj_a"] = np.random.rand(40).tolist()
        weights = {"obj_a": 1.0}
        limits = {"obj_a": [0.0, 1.0]}
        scal, _ = compute_data_array_scalarization(data, weights, copy.deepcopy(limits), "linear")
        data["scalarization"] = list(scal)
        fast = {space.get_unique_hash_string_from_values({p: data[p][r] for p in inputs}): r for r in range(40)}
        best = runner(cfg, data, space, fast, _priors_models(z, prefix), 3, weights, limits)
        assert set(best.keys()) >= set(inputs)
        assert space.get_unique_hash_string_from_values(best) not in fast


# ---- K6 ---------------------------------------------------------------------------------------------------------
def _ks_discrete(sample, values, probs, n):
    """chi-square goodness of fit p-value of a discrete sample"""
    counts = np.array([(sample == v).sum() for v in values], dtype=np.float64)
    assert counts.sum() == n, "values outside the parameter's support"
    keep = np.asarray(probs) > 0
    assert counts[~keep].sum() == 0
    return stats.chisquare(counts[keep], np.asarray(probs)[keep] * n).pvalue


def test_sample_pool_distributions_and_regeneration():
    from hypermapper_b200.space_device import DeviceSpace
    z = load("priors.npz")
    space = space_of(z)
    inputs = space.get_input_parameters()
    objs = space.get_input_parameters_objects()
    ds = DeviceSpace(space)
    N = 200_000
    for use_priors in (False, True):
        out = ds.sample_pool(seed=1234, N=N, use_priors=use_priors, want_raw=True, want32=True, want64=True,
                             prob_weights=[1.0])
        raw = out["raw"].cpu().numpy()                      # [D][N]
        # the encoded matrices of the same pass == encoding the raw values; prob == prior of the raw values
        x32, x64 = ds.encode(np.ascontiguousarray(raw.T), want32=True, want64=True)
        assert np.array_equal(out["x64"].cpu().numpy(), x64.cpu().numpy())
        assert np.array_equal(out["x32"].cpu().numpy(), x32.cpu().numpy())
        assert np.array_equal(out["prob"].cpu().numpy(), ds.prior_prob(np.ascontiguousarray(raw.T), [1.0]).cpu().numpy(),
                              equal_nan=True)
        # any candidate can be regenerated from (seed, index); a pool is the same whether drawn whole or in pieces
        idx = np.array([0, 1, 77, N - 1, 12345], dtype=np.int64)
        assert np.array_equal(ds.sample_rows(1234, use_priors, idx), raw[:, idx].T)
        part = ds.sample_pool(seed=1234, N=1000, use_priors=use_priors, index_base=5000, want_raw=True, want32=False)
        assert np.array_equal(part["raw"].cpu().numpy(), raw[:, 5000:6000])
        other = ds.sample_pool(seed=1235, N=1000, use_priors=use_priors, want_raw=True, want32=False)
        assert not np.array_equal(other["raw"].cpu().numpy(), raw[:, :1000])
        pvals = {}
        for j, p in enumerate(inputs):
            x, o, kind = raw[j], objs[p], space.get_type(p)
            prior = o.prior if use_priors else "uniform"
            if kind == "real":
                lo, hi = o.get_min(), o.get_max()
                assert x.min() >= lo and x.max() <= hi
                if isinstance(prior, list):
                    # np.interp(u, cdf_distribution, param_xs) (space.py:151-153): the cdf is the table itself
                    pvals[p] = stats.kstest(x, lambda v: np.interp(v, o.param_xs, o.cdf_distribution)).pvalue
                elif prior == "custom_gaussian":
                    m, s = o.custom_gaussian_prior_mean, o.custom_gaussian_prior_std
                    a, b = (lo - m) / s, (hi - m) / s
                    pvals[p] = stats.kstest(x, stats.truncnorm(a, b, loc=m, scale=s).cdf).pvalue
                else:
                    al, be = (o.alpha, o.beta) if use_priors else (1.0, 1.0)
                    pvals[p] = stats.kstest((x - lo) / (hi - lo), stats.beta(al, be).cdf).pvalue
            elif kind == "integer":
                vals = list(range(o.get_min(), o.get_max() + 1))
                if prior == "distribution":
                    probs = o.distribution
                elif prior == "uniform" or not use_priors:
                    probs = [1.0 / len(vals)] * len(vals)
                else:   # round(x * size - 0.5 + min) of a Beta draw (space.py:437-453)
                    edges = np.arange(len(vals) + 1) / len(vals)
                    probs = np.diff(stats.beta(o.alpha, o.beta).cdf(edges))
                pvals[p] = _ks_discrete(x, vals, probs, N)
            else:
                vals = list(o.get_values()) if kind == "ordinal" else list(range(o.get_size()))
                if not use_priors:
                    probs = [1.0 / len(vals)] * len(vals)
                else:
                    probs = [o.get_x_probability(v) for v in vals]
                pvals[p] = _ks_discrete(x, np.asarray(vals, dtype=np.float64), probs, N)
        assert min(pvals.values()) > 1e-4, pvals
        # parameters are drawn independently
        c = np.corrcoef(raw)
        assert np.abs(c - np.eye(len(inputs))).max() < 0.02


def test_device_pools_in_local_search_score_like_host_pools():
    """The fused sample+encode pass feeds the same scoring kernels: scores of a device pool == scores of its raw rows
    taken through the host API."""
    from hypermapper_b200 import models, random_scalarizations as rs
    from hypermapper_b200.space_device import device_space_for
    z = load("rf_mixed.npz")
    space = space_of(z)
    objectives = json.loads(str(z["objectives_json"]))
    forests = {o: FrozenForest(z, "forest_%s_" % o) for o in objectives}
    inputs = space.get_input_parameters()
    data = {p: z["data_inputs"][:, j].tolist() for j, p in enumerate(inputs)}
    for o in objectives:
        data[o] = z["data_" + o].tolist()
    weights = {o: float(w) for o, w in zip(objectives, z["weights"])}
    limits = {o: [float(a), float(b)] for o, (a, b) in zip(objectives, z["stored_limits"])}
    ds = device_space_for(space)
    pool = ds.sample_pool(seed=9, N=5000, use_priors=False, want_raw=True, want32=True)
    enc = models.EncodedCandidates(x32=pool["x32"])
    got = rs.acquisition_scores_device("EI", enc, weights, forests, space, "tchebyshev", copy.deepcopy(limits), 7, data,
                                       "random_forest")
    raw_rows = np.ascontiguousarray(pool["raw"].cpu().numpy().T)
    want = rs.acquisition_scores_device("EI", raw_rows, weights, forests, space, "tchebyshev", copy.deepcopy(limits), 7,
                                        data, "random_forest")
    assert np.array_equal(got.cpu().numpy(), want.cpu().numpy())
    confs = ds.rows_to_configurations(raw_rows[:5])
    assert all(isinstance(c["x_cat"], int) and c["x_ord"] in (1, 2, 4, 8, 16, 32, 64, 128) for c in confs)


def test_local_search_with_device_pools():
    """HM_B200_DEVICE_POOLS: the pools are drawn and encoded on the GPU; the optimiser returns an unseen configuration
    whose acquisition value is the minimum over everything it evaluated, and the run is reproducible under
    np.random.seed."""
    import random
    from hypermapper_b200 import local_search as ls, random_scalarizations as rs, utility_functions as uf
    from bench_support.space_lite import SpaceLite
    z = load("local_search.npz")
    config = json.loads(str(z["config_json"]))
    config["local_search_random_points"] = 20000
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
    ls.set_device_pools(True)
    try:
        picks = []
        for rep in range(2):
            np.random.seed(3)
            random.seed(3)
            data_array, best = ls.local_search(4, 20000, space, fast, False, rs.run_acquisition_function, p,
                                               config["scalarization_key"], 1, previous_points=data)
            picks.append(best)
            assert space.get_unique_hash_string_from_values(best) not in fast
            vals, _ = rs.run_acquisition_function(configurations=[best], **p)
            unseen = [v for i, v in enumerate(data_array[config["scalarization_key"]])
                      if space.get_unique_hash_string_from_values({k: data_array[k][i] for k in inputs}) not in fast]
            assert vals[0] == min(unseen)
            # 20000 draws of a 32*7*4*3 = 2688-point space cover it: the pick is the global optimum among unseen points
            everything = [dict(zip(inputs, c)) for c in __import__("itertools").product(
                *[space.get_input_parameters_objects()[k].get_discrete_values() for k in inputs])]
            allv, _ = rs.run_acquisition_function(configurations=everything, **p)
            best_unseen = min(v for c, v in zip(everything, allv)
                              if space.get_unique_hash_string_from_values(c) not in fast)
            assert vals[0] == best_unseen
        assert picks[0] == picks[1]
    finally:
        ls.set_device_pools("auto")
    assert not ls._use_device_pools(10000) and ls._use_device_pools(1 << 20)


def test_pibo_on_a_device_pool_equals_pibo_on_its_rows():
    from hypermapper_b200 import pibo, random_scalarizations as rs
    from hypermapper_b200.space_device import DevicePool, device_space_for
    z = load("priors.npz")
    space = space_of(z)
    inputs = space.get_input_parameters()
    data = {p: z["data_inputs"][:, j].tolist() for j, p in enumerate(inputs)}
    data["obj_a"] = z["data_obj_a"].tolist()
    limits = {"obj_a": [min(data["obj_a"]), max(data["obj_a"])]}
    pool = DevicePool(device_space_for(space), 77, True, 3000)
    args = (space, rs.run_acquisition_function, "EI", {"obj_a": 1.0}, limits, 4, 2.0, _priors_models(z), None, data,
            "random_forest", "linear")
    got = pibo.prior_weighted_scores_device(pool.candidates, *args).cpu().numpy()
    rows = pool.rows(range(3000))
    want, _ = pibo.prior_weighted_acquisition_function(rows, *args)
    assert np.array_equal(got, np.asarray(want), equal_nan=True)


def test_raw_host_scoring_entry_matches_device_path():
    """hm_rf_score_raw_host (raw host rows -> chunked H2D -> hm_encode -> fused forest + acquisition -> D2H) equals the
    device-resident path bit for bit, top-k included, across a chunk boundary; unknown discrete values raise."""
    from hypermapper_b200 import models, random_scalarizations as rs, _lib, forest
    from hypermapper_b200.scoring import HostScorer
    from hypermapper_b200.space_device import device_space_for
    z = load("rf_mixed.npz")
    space = space_of(z)
    objectives = json.loads(str(z["objectives_json"]))
    forests = {o: FrozenForest(z, "forest_%s_" % o) for o in objectives}
    ds = device_space_for(space)
    pool = ds.sample_pool(seed=3, N=5000, use_priors=False, want_raw=True, want32=True)
    raw_rows = np.ascontiguousarray(pool["raw"].cpu().numpy().T)
    dfs = [forest.device_forest(forests[o]) for o in objectives]
    acq = _lib.make_acq_params("EI", "tchebyshev", [0.37, 0.63], [0.4, 0.2], 0.0)
    want = forest.rf_eval(dfs, pool["x32"], acq=acq)["score"]
    wv, wi = forest.topk(want, 20)
    scorer = HostScorer(ds.n_encoded, chunk=2048)             # 3 chunks
    scores = np.empty(5000)
    _, tv, ti = scorer.score_raw(ds, dfs, raw_rows, acq, k=20, scores_out=scores)
    assert np.array_equal(scores, want.cpu().numpy())
    assert np.array_equal(ti, wi.cpu().numpy()) and np.array_equal(tv, wv.cpu().numpy())
    bad = raw_rows.copy()
    bad[4321, 2] = 3.0
    with pytest.raises(ValueError):
        scorer.score_raw(ds, dfs, bad, acq, k=0, scores_out=scores)


def _random_space_config(rng, normalize):
    """a random mix of parameter kinds / sizes / priors (every prior form the device supports)"""
    ip = {}
    n = int(rng.integers(1, 14))
    for j in range(n):
        kind = rng.choice(["real", "integer", "ordinal", "categorical"])
        if kind == "real":
            lo = float(np.round(rng.normal() * 5, 2))
            hi = lo + float(np.round(rng.random() * 10 + 0.5, 2))
            prior = rng.choice(["uniform", "gaussian", "decay", "exponential", "list"])
            p = {"parameter_type": "real", "values": [lo, hi], "parameter_default": lo}
            p["prior"] = [float(v) for v in rng.random(int(rng.integers(2, 12))) + 0.05] if prior == "list" else str(prior)
        elif kind == "integer":
            lo = int(rng.integers(-5, 5))
            hi = lo + int(rng.integers(1, 40))
            p = {"parameter_type": "integer", "values": [lo, hi], "parameter_default": lo}
            if rng.random() < 0.4:
                w = rng.random(hi - lo + 1) + 0.01
                p["prior"] = [float(v) for v in w / w.sum()]
            else:
                p["prior"] = str(rng.choice(["uniform", "gaussian"]))
        elif kind == "ordinal":
            size = int(rng.integers(1, 70))
            vals = np.unique(np.round(np.cumsum(rng.random(size) + 0.01) * (10 ** int(rng.integers(-2, 4))), 6)).tolist()
            if rng.random() < 0.5:
                vals = [int(v) for v in np.unique(np.asarray(vals).astype(np.int64))]
            p = {"parameter_type": "ordinal", "values": vals, "parameter_default": vals[0]}
            if rng.random() < 0.4:
                w = rng.random(len(vals)) + 0.01
                p["prior"] = [float(v) for v in w / w.sum()]
            else:
                p["prior"] = str(rng.choice(["uniform", "gaussian", "decay", "exponential"]))
        else:
            size = int(rng.integers(1, 12))
            p = {"parameter_type": "categorical", "values": ["v%d" % i for i in range(size)], "parameter_default": "v0"}
            if rng.random() < 0.5:
                w = rng.random(size) + 0.01
                p["prior"] = [float(v) for v in w / w.sum()]
        ip["p%d" % j] = p
    return {"optimization_objectives": ["y"], "input_parameters": ip, "normalize_inputs": bool(normalize)}


@pytest.mark.parametrize("seed", range(12))
def test_random_spaces_encode_and_prior_match_host(seed):
    """Random spaces (1-13 parameters of every kind, random sizes, priors and z-score statistics), candidates drawn by
    the host samplers: device encode == host Encoder bit for bit (both kernels: contiguous rows -> tile kernel,
    column-major -> generic kernel); device prior == the product of the parameter objects' get_x_probability."""
    from hypermapper_b200 import _lib
    from hypermapper_b200.encoding import Encoder
    from hypermapper_b200.space_device import DeviceSpace
    from bench_support.space_lite import SpaceLite
    torch = __import__("torch")
    rng = np.random.default_rng(1000 + seed)
    space = SpaceLite(_random_space_config(rng, normalize=seed % 2 == 0))
    inputs = space.get_input_parameters()
    objs = space.get_input_parameters_objects()
    if seed % 2 == 0:
        for p in inputs:
            if space.get_type(p) in ("real", "integer"):
                space.set_parameter_mean(p, float(rng.normal()))
                space.set_parameter_std(p, float(rng.random() + 0.1))
    np.random.seed(seed)
    N = int(rng.integers(1, 3000))
    pools = [space.get_random_configuration(size=N, use_priors=u, return_as_array=True) for u in (False, True)]
    raw = np.concatenate([np.asarray(p, dtype=np.float64).reshape(N, len(inputs)) for p in pools], axis=0)
    ds = DeviceSpace(space)
    want = Encoder(space).encode(raw)
    x32, x64 = ds.encode(raw, want32=True, want64=True)
    assert np.array_equal(x64.cpu().numpy(), want) and np.array_equal(x32.cpu().numpy(), want.astype(np.float32))
    # column-major raw input -> the generic kernel
    rd = torch.from_numpy(np.ascontiguousarray(raw.T)).cuda()
    g64 = torch.empty_like(x64)
    bad = torch.zeros((1,), dtype=torch.int32, device="cuda")
    rc = _lib.lib().hm_encode(ds.handle, _lib.dptr(rd), 1, raw.shape[0], raw.shape[0], None, 0, _lib.dptr(g64),
                              raw.shape[0], _lib.dptr(bad), _lib.stream_ptr())
    _lib.check(rc, "hm_encode")
    assert int(bad.item()) == 0 and np.array_equal(g64.cpu().numpy(), want)
    # priors
    got = ds.prior_prob(raw, [1.0]).cpu().numpy()
    host = np.ones(len(raw))
    for j, p in enumerate(inputs):
        kind = space.get_type(p)
        col = raw[:, j]
        if kind == "real":
            px = np.array([objs[p].get_x_probability(float(v)) for v in col], dtype=np.float64)
        elif kind == "ordinal":
            vals = objs[p].get_values()
            lut = {float(v): v for v in vals}
            px = np.array([objs[p].get_x_probability(lut[float(v)]) for v in col], dtype=np.float64)
        else:
            px = np.array([objs[p].get_x_probability(int(v)) for v in col], dtype=np.float64)
        host = host * px
    close(got, host, rel=1e-9)
    gp = ds.prior_prob(rd, [1.0], col_major=True).cpu().numpy()
    assert np.array_equal(gp, got, equal_nan=True)


def test_device_neighbourhoods_equal_get_neighbors():
    """hm_ls_neighbors against get_neighbors row for row (order included) on every window position of every ordinal,
    and hm_ls_move against a host argmin with the reference's tie-break."""
    import ctypes
    import torch
    from hypermapper_b200 import _lib
    from hypermapper_b200.local_search import get_neighbors, neighbor_plan
    from hypermapper_b200.space_device import device_space_for
    from bench_support.space_lite import SpaceLite
    cfg = {"optimization_objectives": ["y"], "input_parameters": {
        "big": {"parameter_type": "ordinal", "values": [1, 2, 4, 8, 16, 32, 64, 128, 256], "parameter_default": 1},
        "cat": {"parameter_type": "categorical", "values": ["a", "b", "c"], "parameter_default": "a"},
        "five": {"parameter_type": "ordinal", "values": [0.5, 1.5, 2.5, 3.5, 4.5], "parameter_default": 0.5},
        "six": {"parameter_type": "ordinal", "values": [10, 20, 30, 40, 50, 60], "parameter_default": 10},
        "one": {"parameter_type": "categorical", "values": ["z"], "parameter_default": "z"}}}
    space = SpaceLite(cfg)
    inputs = space.get_input_parameters()
    objs = space.get_input_parameters_objects()
    plan = neighbor_plan(space)
    K = len(plan)
    ds = device_space_for(space)
    rng = np.random.default_rng(0)
    starts = []
    for i in range(9):                                     # every index of the 9-value ordinal, others random
        starts.append({"big": objs["big"].get_values()[i], "cat": int(rng.integers(0, 3)),
                       "five": objs["five"].get_values()[int(rng.integers(0, 5))],
                       "six": objs["six"].get_values()[i % 6], "one": 0})
    S = len(starts)
    cur = torch.from_numpy(np.array([[float(c[p]) for p in inputs] for c in starts])).cuda()
    live = torch.arange(S, dtype=torch.int32, device="cuda")
    rows = torch.empty((S * K, len(inputs)), dtype=torch.float64, device="cuda")
    L = _lib.lib()
    _lib.check(L.hm_ls_neighbors(ds.handle, _lib.dptr(torch.from_numpy(plan).cuda()), K, _lib.dptr(cur), _lib.dptr(live), S,
                                 _lib.dptr(rows), _lib.stream_ptr()), "hm_ls_neighbors")
    got = rows.cpu().numpy().reshape(S, K, len(inputs))
    for s, c in enumerate(starts):
        want = np.array([[float(v) for v in r] for r in get_neighbors(c, space)])
        assert want.shape == (K, len(inputs)) and np.array_equal(got[s], want), s
    # moves: ties -> first index, NaN first, -0.0 == +0.0
    scores = rng.integers(0, 3, (S, K)).astype(np.float64)
    scores[1, 5] = np.nan
    scores[2, 3] = -0.0
    scores[2, 1] = 0.0
    scores[2][scores[2] > 0] += 1
    sd = torch.from_numpy(scores.reshape(-1)).cuda()
    conv = torch.zeros((S,), dtype=torch.int32, device="cuda")
    bq = torch.zeros((S,), dtype=torch.int32, device="cuda")
    before = cur.clone()
    _lib.check(L.hm_ls_move(_lib.dptr(sd), _lib.dptr(rows), K, len(inputs), _lib.dptr(live), S, _lib.dptr(cur),
                            _lib.dptr(conv), _lib.dptr(bq), _lib.stream_ptr()), "hm_ls_move")
    for s in range(S):
        v = scores[s]
        q = min(range(K), key=lambda i: (0, 0.0) if v[i] != v[i] else (1, v[i] + 0.0))
        assert int(bq[s]) == q, (s, int(bq[s]), q)
        same = np.array_equal(got[s, q], before[s].cpu().numpy())
        assert int(conv[s]) == int(same)
        assert np.array_equal(cur[s].cpu().numpy(), got[s, q])
