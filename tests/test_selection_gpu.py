"""Selection identity ("ties are the norm"): GPU-computed scores -> stable top-k -> the indices the
REFERENCE's get_min_configurations picks on ITS OWN values (utility_functions.py:295-316), with the near-tied finalists
re-scored on the host by the reference's arithmetic (hypermapper_b200/exact.py).  Covers the golden vectors of the
reference, the bench's own C3 / C5 workloads at 2^20 candidates, and recorded local_search() runs of the reference on
real / integer / mixed / discrete spaces."""
import json

import numpy as np
import pytest

import ls_replay
from test_api_gpu import FrozenForest, load, setup_case

pytestmark = pytest.mark.gpu


def _exact_fn(acq, scal, models, X32, weights, limits, iteration, data):
    """idx -> the reference's exact values of pool rows idx (leaves from the GPU, arithmetic on the host)"""
    import torch
    from hypermapper_b200 import exact, forest
    objectives = list(models.keys())
    forests = [forest.device_forest(models[o]) for o in objectives]

    def f(idx):
        sub = X32[:, torch.as_tensor(np.asarray(idx, dtype=np.int64), device=X32.device)].contiguous()
        leaves = [lv.cpu().numpy() for lv in forest.rf_eval(forests, sub, want_leaves=True)["leaves"]]
        return exact.reference_rf_scores_from_leaves(acq, scal, models, leaves, weights, limits, iteration, data)
    return f, forests


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
@pytest.mark.parametrize("acq", ["EI", "UCB", "TS"])
@pytest.mark.parametrize("scal", ["linear", "tchebyshev", "modified_tchebyshev"])
def test_gpu_scores_to_topk_equals_reference_selection(name, acq, scal):
    """rf_eval(golden candidates) -> top-k == get_min_configurations of the reference on the reference's values"""
    import torch
    from hypermapper_b200 import exact, forest, _lib
    from hypermapper_b200.random_scalarizations import acquisition_constants
    from oracle import oracle
    z, space, objectives, models, bufferx, data, weights, limits = setup_case(name)
    it = int(z["iteration_number"])
    X32 = torch.from_numpy(np.ascontiguousarray(z["encoded"].astype(np.float32).T)).cuda()
    f, forests = _exact_fn(acq, scal, models, X32, weights, limits, it, data)
    w, fmin, beta = acquisition_constants(acq, scal, objectives, weights, limits, data, it)
    scores = forest.rf_eval(forests, X32, acq=_lib.make_acq_params(acq, scal, w, fmin, beta))["score"]
    ref_vals = z["acq_%s_%s" % (acq, scal)]
    for k in (1, 20, 64):
        want = oracle.get_min_indices(ref_vals, k)
        if acq == "EI" and scal == "tchebyshev":
            assert np.array_equal(want, z["topk_idx_%d" % k])          # the reference's own get_min_configurations output
        got = exact.exact_min_indices(scores, k, f)
        assert np.array_equal(got, want), (k, got, want)
    # the host re-score itself is the reference value, bit for bit
    some = np.arange(0, len(ref_vals), 7)
    assert np.array_equal(f(some), ref_vals[some])


def test_near_tie_mechanism_reorders_with_exact_values():
    """distinct device values closer than 1e-9 relative are ordered by the exact values; far ones and bit-equal ones
    (true ties -> lowest index) are left alone; NaN stays first."""
    import torch
    from hypermapper_b200 import exact
    rng = np.random.default_rng(3)
    n = 200_000
    v = rng.standard_normal(n) * 10 - 50
    v[1234] = -200.0
    v[777] = -200.0 * (1 - 3e-13)           # near tie, device says 1234 < 777
    v[90_000] = -200.0                       # bit-equal to 1234: true tie, lowest index first
    v[5] = np.nan
    v[150_000] = -199.0
    truth = {1234: -200.0, 90_000: -200.0, 777: -200.0 * (1 + 1e-15)}        # "reference" says 777 is the smallest
    calls = []

    def f(idx):
        calls.append(list(idx))
        return np.array([truth[int(i)] for i in idx])
    vals = torch.from_numpy(v).cuda()
    got = exact.exact_min_indices(vals, 4, f)
    assert got.tolist() == [5, 777, 1234, 90_000]
    assert sorted(calls[0]) == [777, 1234]                    # one representative per distinct value
    got = exact.exact_min_indices(vals, 5, f)
    assert got.tolist() == [5, 777, 1234, 90_000, 150_000]
    # no near tie around the boundary -> no re-score at all
    calls.clear()
    assert exact.exact_min_indices(torch.from_numpy(rng.standard_normal(5000)).cuda(), 20, f).shape == (20,)
    assert calls == []


@pytest.mark.parametrize("which", ["c3", "c5rf"])
def test_bench_workload_top20_equals_oracle_selection(which):
    """the bench's own fitted 2 x 256-tree (C3) / 256-tree (C5) forests, a seeded 2^20-candidate pool: the GPU's top-20
    == the oracle's get_min_configurations on the oracle's values; leaves / means bit-exact on a sample."""
    import torch
    from hypermapper_b200 import exact, forest
    from bench_support import workloads
    from oracle import oracle
    if which == "c3":
        wl = workloads.RFWorkload(workloads.scenario_c3(), n_train=2000, seed=5)
    else:
        wl = workloads.RFWorkload(workloads.scenario_c5(), n_train=4000, seed=9)
    oracle.use_all_host_threads()
    N = 1 << 20
    Xh = wl.candidates_encoded(N, seed=77)
    X32 = torch.from_numpy(Xh).cuda()
    f, forests = _exact_fn(wl.acquisition, wl.scalarization, wl.models, X32, wl.objective_weights, wl.objective_limits,
                           wl.iteration_number, wl.data_array)
    out = forest.rf_eval(forests, X32, acq=wl.acq_params())
    fas = {o: oracle.forest_arrays_from_model(m) for o, m in wl.models.items()}
    Xrows = np.ascontiguousarray(Xh.T)
    means, variances = {}, {}
    for o, fa in fas.items():
        means[o], variances[o], _ = oracle.rf_mean_var_omp(fa, Xrows)
    want_vals = oracle.EI(means, variances, wl.data_array, wl.objective_weights, wl.scalarization, wl.objective_limits)
    want = oracle.get_min_indices_fast(want_vals, 20)
    got = exact.exact_min_indices(out["score"], 20, f)
    assert np.array_equal(got, want), (got, want)
    s = out["score"].cpu().numpy()
    err = np.abs(s - want_vals) / np.maximum(np.abs(want_vals), 1e-300)
    err[s == want_vals] = 0
    assert err.max() <= 1e-5
    mv = forest.rf_eval(forests, X32[:, :65536].contiguous(), want_mean=True)
    for m, o in enumerate(wl.objectives):
        assert np.array_equal(mv["mean"][m].cpu().numpy(), means[o][:65536])


@pytest.mark.parametrize("tag", ls_replay.TAGS)
def test_local_search_replays_reference_runs_on_gpu(tag):
    """the reference's recorded local_search() runs (real / integer / mixed / discrete spaces) through
    run_acquisition_function on the GPU: every evaluated configuration in the same order, values within 1e-5, and the
    same pick."""
    from hypermapper_b200 import random_scalarizations as rs
    z = ls_replay.load()
    out = ls_replay.replay(z, tag, rs.run_acquisition_function, FrozenForest)
    ls_replay.check(*out, rel=1e-5)


def test_final_pick_when_the_best_of_a_pool_is_already_evaluated():
    """more than 256 of the best pool entries were evaluated before (late in a run on a small discrete space): the
    search goes past the k <= 1024 limit of hm_topk instead of raising (local_search.py:789-807 keeps deleting)."""
    from hypermapper_b200 import local_search as ls, random_scalarizations as rs, utility_functions as uf
    from bench_support.space_lite import SpaceLite
    import copy
    z = load("local_search.npz")
    config = json.loads(str(z["config_json"]))
    space = SpaceLite(config)
    inputs = space.get_input_parameters()
    models = {"Value": FrozenForest(z, "forest_Value_")}
    data = {k: [int(v) for v in z["data_inputs"][:, j]] for j, k in enumerate(inputs)}
    data["Value"] = z["data_Value"].tolist()
    weights = {"Value": 1.0}
    limits = {"Value": [min(data["Value"]), max(data["Value"])]}
    scal, _ = uf.compute_data_array_scalarization(data, weights, copy.deepcopy(limits), "tchebyshev")
    data[config["scalarization_key"]] = scal.tolist()
    p = {"regression_models": models, "iteration_number": 3, "data_array": data, "classification_model": None,
         "param_space": space, "objective_weights": weights, "model_type": "random_forest",
         "objective_limits": limits, "acquisition_function": "EI", "scalarization_method": "tchebyshev",
         "number_of_cpus": 1}
    # everything except a thin slice of the space counts as evaluated
    everything = uf.dict_of_lists_to_list_of_dicts(space.get_space())
    fast = {space.get_unique_hash_string_from_values(c): 1 for c in everything if not (c["a"] == 32 and c["d"] == 2)}
    np.random.seed(1)
    data_array, best = ls.local_search(4, 2500, space, fast, False, rs.run_acquisition_function, p,
                                       config["scalarization_key"], 1, previous_points=data)
    assert best["a"] == 32 and best["d"] == 2
    assert space.get_unique_hash_string_from_values(best) not in fast


@pytest.mark.parametrize("case", ["c3_packed", "f32_ties", "nan_scores", "overflow", "short_pool", "k1024"])
def test_fused_score_topk_equals_eval_then_topk(case):
    """hm_rf_score_topk(_packed) -- threshold filter in the forest kernel's epilogue, no score array -- returns exactly
    what hm_rf_eval + hm_topk returns: tie classes larger than k, NaN scores (np.argmin order: NaN first), a pool whose
    scores keep improving (list overflow -> fall-back), pools shorter than the prefix, k = 1024."""
    import torch
    from hypermapper_b200 import forest, _lib
    from test_packed_gpu import _make, _raw
    from test_rf_gpu import _synthetic_forest
    rng = np.random.default_rng(11)
    k, feas = 20, None
    if case in ("c3_packed", "k1024"):
        space, ds = _make("c3")
        N = 700_001
        raw = _raw(space, N, rng)
        X = ds.encode_packed(raw)
        x32, _ = ds.encode(raw)
        fs = [_synthetic_forest(rng, 24, ds.n_encoded, 5, 9) for _ in range(2)]       # shallow: huge tie classes
        dfs = [forest.DeviceForest(forest.tables_from_arrays(**f)) for f in fs]
        forest.attach_packing(dfs, ds)
        k = 1024 if case == "k1024" else 20
        plain = lambda acq, fe: forest.rf_eval_packed(dfs, X, acq=acq, feas=fe)["score"]
    else:
        D = 5
        N = 1000 if case == "short_pool" else 500_003
        Xh = rng.random((N, D))
        if case == "f32_ties":
            Xh = np.round(Xh * 3) / 3                                                  # 4^5 distinct candidates
        X = torch.from_numpy(np.ascontiguousarray(Xh.astype(np.float32).T)).cuda()
        fs = [_synthetic_forest(rng, 12, D, 3, 8) for _ in range(2)]
        dfs = [forest.DeviceForest(forest.tables_from_arrays(**f)) for f in fs]
        if case == "nan_scores":
            fe = np.ones(N)
            fe[[5, 70_000, 70_001, 400_000, N - 1]] = np.nan                           # NaN scores inside and beyond the prefix
            feas = torch.from_numpy(fe).cuda()
        if case == "overflow":
            feas = torch.linspace(1.0, 50.0, N, dtype=torch.float64).cuda()             # -EI * feas: ever better scores
        plain = lambda acq, fe: forest.rf_eval(dfs, X, acq=acq, feas=fe)["score"]
    for kind, scal in (("EI", "tchebyshev"), ("TS", "linear"), ("UCB", "modified_tchebyshev")):
        acq = _lib.make_acq_params(kind, scal, [0.3, 0.7], [0.1, 0.2], 0.5)
        want_v, want_i = forest.topk(plain(acq, feas), k)
        got_v, got_i = forest.score_topk(dfs, X, acq, k, feas=feas)
        torch.cuda.synchronize()
        assert torch.equal(got_i, want_i), (case, kind)
        assert torch.equal(got_v.view(torch.int64), want_v.view(torch.int64)), (case, kind)
