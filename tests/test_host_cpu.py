"""CPU tests of the host-side logic (no GPU, no compute calls): encoding, scalarisation constants, neighbour
generation, selection order, the C-ABI surface, and the world_size-2 top-k merge over gloo."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def space_of(z, key="config_json"):
    from bench_support.space_lite import SpaceLite
    return SpaceLite(json.loads(str(z[key])))


def typed_rows(params, arr):
    rows = []
    for row in arr:
        t = []
        for p, v in zip(params, row):
            if p["type"] in ("integer", "categorical") or (p["type"] == "ordinal" and all(float(x).is_integer() for x in p["values"])):
                t.append(int(v))
            else:
                t.append(float(v))
        rows.append(tuple(t))
    return rows


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
def test_encoder_matches_reference_preprocess(name):
    from hypermapper_b200.encoding import Encoder
    z = load(name)
    space = space_of(z)
    enc = Encoder(space)
    params = json.loads(str(z["params_json"]))
    rows = typed_rows(params, z["candidates"])
    assert np.array_equal(enc.encode(rows).T, z["encoded"])                     # list of tuples
    assert np.array_equal(enc.encode(z["candidates"]).T, z["encoded"])          # 2-D float array
    cfgs = [dict(zip(space.get_input_parameters(), r)) for r in rows]
    assert np.array_equal(enc.encode_configurations(cfgs).T, z["encoded"])      # list of dicts
    assert enc.n_encoded == z["encoded"].shape[1]


def test_encoder_rejects_unknown_values():
    from hypermapper_b200.encoding import Encoder
    z = load("rf_mixed.npz")
    enc = Encoder(space_of(z))
    bad = z["candidates"][:3].copy()
    bad[1, 2] = 3.0                 # not one of the ordinal's values
    with pytest.raises(ValueError):
        enc.encode(bad)
    bad = z["candidates"][:3].copy()
    bad[0, 3] = 7                   # categorical index out of range
    with pytest.raises(ValueError):
        enc.encode(bad)


def test_encoder_normalize_inputs():
    from hypermapper_b200.encoding import Encoder
    from bench_support.space_lite import SpaceLite
    from oracle import oracle
    z = load("rf_mixed.npz")
    cfg = json.loads(str(z["config_json"]))
    cfg["normalize_inputs"] = True
    space = SpaceLite(cfg)
    params = json.loads(str(z["params_json"]))
    means, stds = {}, {}
    for j, p in enumerate(params):
        means[p["name"]] = float(np.mean(z["candidates"][:, j]))
        stds[p["name"]] = float(np.std(z["candidates"][:, j]))
        space.set_parameter_mean(p["name"], means[p["name"]])
        space.set_parameter_std(p["name"], stds[p["name"]])
    rows = typed_rows(params, z["candidates"])
    cols = {p["name"]: [r[j] for r in rows] for j, p in enumerate(params)}
    want, _ = oracle.encode(cols, params, normalize_inputs=True, means=means, stds=stds)
    assert np.array_equal(Encoder(space).encode(rows).T, want)


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
@pytest.mark.parametrize("scal", ["linear", "tchebyshev", "modified_tchebyshev"])
def test_scalarization_and_constants_match_reference(name, scal):
    from hypermapper_b200 import utility_functions as uf
    from hypermapper_b200.random_scalarizations import acquisition_constants
    from oracle import oracle
    z = load(name)
    objectives = json.loads(str(z["objectives_json"]))
    data = {o: z["data_" + o].tolist() for o in objectives}
    weights = {o: float(w) for o, w in zip(objectives, z["weights"])}
    limits = {o: [float(a), float(b)] for o, (a, b) in zip(objectives, z["stored_limits"])}
    s, lim = uf.compute_data_array_scalarization(data, weights, limits, scal)
    assert np.array_equal(s, z["data_scalarization_" + scal])
    assert np.array_equal(np.array([lim[o] for o in objectives]), z["data_scalarization_limits_" + scal])
    w, fmin, beta = acquisition_constants("EI", scal, objectives, weights, limits, data, 7)
    assert fmin == [oracle.ei_f_min(data, o, lim) for o in objectives]
    rw = oracle.reciprocate_weights(weights)
    assert w == [(rw if scal == "modified_tchebyshev" else weights)[o] for o in objectives]
    _, _, beta = acquisition_constants("UCB", scal, objectives, weights, limits, data, 7)
    assert beta == float(np.sqrt(0.125 * np.log(2 * 7 + 1)))


def test_min_indices_host_order():
    from hypermapper_b200 import utility_functions as uf
    from oracle import oracle
    rng = np.random.default_rng(0)
    v = np.round(rng.standard_normal(3000), 1)
    v[[17, 400]] = np.nan
    v[[3, 9]] = -0.0
    for k in (1, 7, 100, 5000):
        assert np.array_equal(uf.min_indices(v.tolist(), k), oracle.get_min_indices(v, k))
    d = {"x": list(range(10)), "s": [3, 1, 2, 1, 5, 1, 0, 9, 0, 4], "Valid": [1, 0, 1, 1, 0, 1, 0, 1, 1, 1]}
    best = uf.get_min_feasible_configurations(d, 3, "s", "Valid")
    assert best["x"] == [8, 3, 5]
    best = uf.get_min_feasible_configurations(d, 9, "s", "Valid")       # not enough feasible: all feasible + best infeasible
    assert best["x"] == [0, 2, 3, 5, 7, 8, 9, 6, 1]


@pytest.mark.parametrize("tag", ["discrete", "mixed"])
def test_get_neighbors_matches_reference(tag):
    """same enumeration order and the same truncnorm draws as local_search.get_neighbors of the reference"""
    from hypermapper_b200 import local_search as ls
    z = load("neighbors.npz")
    space = space_of(z, "config_json_" + tag)
    params = json.loads(str(z["params_json_" + tag]))
    inputs = space.get_input_parameters()
    starts = typed_rows(params, z["starts_" + tag])
    for r, row in enumerate(starts):
        np.random.seed(100 + r)
        nb = ls.get_neighbors(dict(zip(inputs, row)), space)
        got = np.array([[float(v) for v in x] for x in nb], dtype=np.float64)
        assert np.array_equal(got, z["neighbors_%s_%d" % (tag, r)])


def test_space_lite_accessors():
    z = load("rf_mixed.npz")
    space = space_of(z)
    assert space.get_input_parameters() == ["x_real", "x_int", "x_ord", "x_cat", "x_ordf", "x_cat2"]
    assert space.get_input_categorical_parameters(space.get_input_parameters()) == ["x_cat", "x_cat2"]
    assert space.get_type("x_ord") == "ordinal" and space.get_space_size() == float("inf")
    np.random.seed(0)
    arr = space.get_random_configuration(size=50, use_priors=False, return_as_array=True)
    assert arr.shape == (50, 6) and arr.dtype == object
    conf = space.get_default_or_random_configuration()
    assert conf["x_cat"] == 0 and conf["x_ordf"] == 1.5
    assert space.get_unique_hash_string_from_values({"x_real": 1.5, "x_int": 3}) == "1.5_3_____"


def test_library_exports_every_declared_symbol():
    from hypermapper_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "hm_b200.h")).read()
    declared = set(re.findall(r"\b(hm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.SO_PATH)          # loads without a GPU
    for sym in declared:
        assert getattr(L, sym) is not None
    L.hm_version.restype = ctypes.c_int
    assert L.hm_version() >= 100
    L.hm_topk_workspace_bytes.restype = ctypes.c_int64
    assert L.hm_topk_workspace_bytes(20) > 0


def test_product_never_imports_the_oracle():
    """the oracle is the checker: nothing under hypermapper_b200/ may import it"""
    pkg = os.path.join(ROOT, "hypermapper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "hm_oracle" not in src, f


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from hypermapper_b200 import _lib, compute_pareto
    with pytest.raises(_lib.HmError):
        compute_pareto.sequential_is_pareto_efficient(np.random.rand(10, 2))


def test_shard_bounds_cover_the_pool():
    from hypermapper_b200.distributed import shard_bounds
    for n, g in ((10, 3), (64, 8), (7, 8), (1 << 26, 8)):
        b = [shard_bounds(n, r, g) for r in range(g)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(g - 1))


GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %r)
from hypermapper_b200.distributed import allgather_topk, shard_bounds, merge_topk_host
from oracle import oracle
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(11)
vals = np.round(rng.standard_normal(50001), 2)       # heavy ties across the shard boundary
vals[[5, 40000]] = np.nan
k = 20
lo, hi = shard_bounds(len(vals), rank, world)
local = oracle.get_min_indices_fast(vals[lo:hi], k)
lv = torch.full((k,), float("inf"), dtype=torch.float64); li = torch.full((k,), -1, dtype=torch.int64)
lv[:len(local)] = torch.from_numpy(vals[lo:hi][local]); li[:len(local)] = torch.from_numpy(local + lo)
gv, gi = allgather_topk(lv, li, k)
want = oracle.get_min_indices_fast(vals, k)
assert np.array_equal(gi.numpy(), want), (rank, gi.numpy(), want)
gathered = [None] * world
dist.all_gather_object(gathered, gi.tolist())
assert all(g == gathered[0] for g in gathered)        # identical on every rank
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_topk_merge_over_gloo(tmp_path):
    """N > 1 path on CPU: contiguous shards, per-rank stable top-k with global indices, all-gather, merge ->
    every rank holds the single-process top-k (ties resolved by lowest global index)."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


GLOO_PARETO_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %r)
from hypermapper_b200.distributed import pareto_mask_sharded, shard_bounds
from oracle import oracle
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)

def cpu_filter(costs, pass2):           # the oracle stands in for the K4 kernels on this GPU-less box
    c = costs.numpy()
    if len(c) == 0:
        return torch.zeros((0,), dtype=torch.uint8)
    return torch.from_numpy((oracle.pareto(c) if pass2 else oracle.pareto_pass1(c)).astype(np.uint8))

rng = np.random.default_rng(3)
cases = {"uniform3": rng.random((6001, 3)), "ties3": np.floor(rng.random((5000, 3)) * 12) / 12,
         "ties2": np.floor(rng.random((4000, 2)) * 40) / 40, "quirk": np.array([[1., 5., 9.], [1., 4., 10.], [2., 2., 2.]]),
         "dups": np.repeat(rng.random((300, 3)), 3, axis=0)}
for name, c in cases.items():
    lo, hi = shard_bounds(len(c), rank, world)
    got = pareto_mask_sharded(torch.from_numpy(np.ascontiguousarray(c[lo:hi])), filter_fn=cpu_filter).numpy().astype(bool)
    want = oracle.pareto(c)[lo:hi]
    assert np.array_equal(got, want), (name, rank, got.sum(), want.sum())
import hypermapper_b200.distributed as D
D._PARETO_EXCHANGE_ROWS = 16            # force the overflow round: an anti-correlated front keeps every point
x = rng.random(3000); front = np.stack([x, 1 - x], 1)
lo, hi = shard_bounds(len(front), rank, world)
got = pareto_mask_sharded(torch.from_numpy(np.ascontiguousarray(front[lo:hi])), filter_fn=cpu_filter).numpy().astype(bool)
assert np.array_equal(got, oracle.pareto(front)[lo:hi]) and got.sum() > 1000
# NaN costs: no pre-filter is valid; every rank falls back to the whole matrix in index order
bad = rng.random((301, 3)); bad[5, 1] = np.nan; bad[200, 0] = np.nan; bad[201] = bad[7]
lo, hi = shard_bounds(len(bad), rank, world)
got = pareto_mask_sharded(torch.from_numpy(np.ascontiguousarray(bad[lo:hi])), filter_fn=cpu_filter).numpy().astype(bool)
assert np.array_equal(got, oracle.pareto(bad)[lo:hi]), ("nan", rank)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_sharded_pareto_over_gloo(tmp_path):
    """SURVEY 8(e) row 2 on CPU: contiguous row blocks, local pass 1, all-gather of the survivors, full filter on the
    union -> every rank holds its slice of the single-process mask (pass 2's index-order dependence included)."""
    script = tmp_path / "worker_pareto.py"
    script.write_text(GLOO_PARETO_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_install_rebinds_reference_entry_points():
    """with the real reference importable (build container only): install() rebinds every hot-path name"""
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("reference tree not present on this box")
    ns = ref_harness.load()
    import hypermapper_b200
    from hypermapper_b200 import models as b_models, random_scalarizations as b_rs, local_search as b_ls, compute_pareto as b_cp
    saved = {}
    targets = [(ns.models, ["compute_model_mean_and_uncertainty", "compute_rf_prediction_mean_and_uncertainty",
                            "compute_gp_prediction_mean_and_uncertainty", "model_prediction", "model_probabilities",
                            "local_search"]),
               (ns.models.RFModel, ["get_leaves_per_sample", "compute_rf_prediction", "compute_rf_prediction_variance"]),
               (ns.random_scalarizations, ["run_acquisition_function", "EI", "ucb", "thompson_sampling", "random_scalarizations"]),
               (ns.local_search, ["local_search", "get_neighbors"]),
               (ns.compute_pareto, ["sequential_is_pareto_efficient", "sequential_is_pareto_efficient_dumb",
                                    "parallel_is_pareto_efficient", "compute_pareto"]),
               (ns.prior_optimization, ["compute_probability_from_prior", "estimate_prior_limits",
                                        "compute_probability_from_model", "compute_EI_from_posteriors",
                                        "prior_guided_optimization", "local_search"]),
               (ns.pibo, ["prior_weighted_acquisition_function", "run_pibo", "compute_probability_from_prior",
                          "run_acquisition_function", "local_search"])]
    for tgt, names in targets:
        for n in names:
            saved[(tgt, n)] = getattr(tgt, n)
    try:
        done = hypermapper_b200.install()
        assert len(done) >= 20
        assert ns.models.compute_model_mean_and_uncertainty is b_models.compute_model_mean_and_uncertainty
        assert ns.random_scalarizations.run_acquisition_function is b_rs.run_acquisition_function
        assert ns.local_search.local_search is b_ls.local_search
        assert ns.compute_pareto.sequential_is_pareto_efficient is b_cp.sequential_is_pareto_efficient
        assert ns.models.RFModel.get_leaves_per_sample is b_models.rf_get_leaves_per_sample
        from hypermapper_b200 import pibo as b_pibo, prior_optimization as b_po
        assert ns.pibo.prior_weighted_acquisition_function is b_pibo.prior_weighted_acquisition_function
        assert ns.prior_optimization.compute_EI_from_posteriors is b_po.compute_EI_from_posteriors
        # signatures keep the reference's positional parameter names
        import inspect
        for (tgt, n), ref_fn in saved.items():
            if tgt is ns.models.RFModel:
                continue
            ref_params = list(inspect.signature(ref_fn).parameters)
            new_params = list(inspect.signature(getattr(tgt, n)).parameters)
            assert new_params[: len(ref_params)] == ref_params, (n, ref_params, new_params)
    finally:
        for (tgt, n), fn in saved.items():
            setattr(tgt, n, fn)


# ---- search-space descriptors (space_device.describe_space) -----------------------------------------------------
def test_space_lite_priors_match_reference_golden():
    """SpaceLite's per-parameter get_x_probability (the tables the device uses are built by calling it) against
    compute_probability_from_prior of the unmodified reference (tests/golden/priors.npz)."""
    z = load("priors.npz")
    space = space_of(z)
    inputs = space.get_input_parameters()
    objs = space.get_input_parameters_objects()
    kinds = [space.get_type(p) for p in inputs]
    got = []
    for row in z["candidates"]:
        prob = 1
        for p, k, v in zip(inputs, kinds, row):
            x = float(v) if k == "real" else (int(v) if k != "ordinal" else
                                             objs[p].get_values()[[float(t) for t in objs[p].get_values()].index(v)])
            prob *= objs[p].get_x_probability(x) ** 1.0
        got.append(prob)
    assert np.array_equal(np.array(got, dtype=np.float64), z["prior_w1"])


def test_describe_space_layout_and_tables():
    from hypermapper_b200 import space_device as sd
    z = load("priors.npz")
    space = space_of(z)
    descs, tables, n_enc = sd.describe_space(space)
    inputs = space.get_input_parameters()
    objs = space.get_input_parameters_objects()
    assert n_enc == 8 + 3 + 2 and len(descs) == len(inputs) == 10
    by = {p: descs[j] for j, p in enumerate(inputs)}
    # encoded column order: non-categoricals in input order, then the one-hot blocks (models.py:957-984)
    assert [by[p].out_col for p in inputs] == [0, 1, 2, 3, 4, 5, 6, 7, 8, 11]
    assert by["r_gauss"].prior_kind == sd.PRIOR_BETA and (by["r_gauss"].a, by["r_gauss"].b) == (3.0, 3.0)
    assert by["r_cg"].prior_kind == sd.PRIOR_GAUSSIAN and (by["r_cg"].a, by["r_cg"].b) == (0.3, 0.15)
    d = by["r_list"]
    assert d.prior_kind == sd.PRIOR_INTERP and d.n_prior == 10 and d.n_cdf == 10000
    assert np.array_equal(tables[d.prior_off + 10: d.prior_off + 20], np.asarray(objs["r_list"].prior))
    assert tables[d.cdf_off + d.n_cdf - 1] == 1.0
    for p in ("i_exp", "i_list", "o_gauss", "o_list", "c_list", "c_uni"):
        d, o = by[p], objs[p]
        assert d.prior_kind == sd.PRIOR_TABLE
        vals = (range(o.get_min(), o.get_max() + 1) if space.get_type(p) == "integer" else
                (o.get_values() if space.get_type(p) == "ordinal" else range(o.get_size())))
        assert np.array_equal(tables[d.prior_off: d.prior_off + d.n_prior],
                              np.array([o.get_x_probability(v) for v in vals], dtype=np.float64))
        assert d.ordinal_beta == (1 if p in ("i_exp", "o_gauss") else 0)
    d = by["o_gauss"]
    assert np.array_equal(tables[d.values_off: d.values_off + 8], [1, 2, 4, 8, 16, 32, 64, 128])
    assert np.array_equal(tables[d.values_off + 8: d.values_off + 16], [k / 8 for k in range(8)])
    assert ctypes.sizeof(sd.ParamDesc) == 8 * 4 + 3 * 8 + 7 * 8


def test_describe_space_same_for_reference_space():
    """The descriptors built from the real reference Space equal those built from SpaceLite."""
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("reference tree not present")
    from hypermapper_b200 import space_device as sd
    ns = ref_harness.load()
    z = load("priors.npz")
    cfg = json.loads(str(z["config_json"]))
    ref_space = ns.space.Space(cfg)
    d0, t0, n0 = sd.describe_space(ref_space)
    d1, t1, n1 = sd.describe_space(space_of(z))
    assert n0 == n1 and np.array_equal(t0, t1)
    assert bytes(d0) == bytes(d1)


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference` runs without a GPU (it times the CPU port) and prints ONE JSON line with the keys the
    driver reads; the default arm refuses to run without a CUDA device instead of falling back."""
    env = dict(os.environ, HM_BENCH_CACHE=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "2", "--warmup", "1", "--cpu-sample", "20000"], env=env, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "impl", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    import torch
    if not torch.cuda.is_available():
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1", "--steps", "1",
                              "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode != 0 and "CUDA" in (out.stderr + out.stdout)


def test_neighbor_plan_enumerates_like_get_neighbors():
    """local_search.neighbor_plan (what hm_ls_neighbors follows on the device) against get_neighbors on the discrete
    golden space: same number of neighbours for every configuration, same order, same rows."""
    from hypermapper_b200.local_search import get_neighbors, neighbor_plan
    z = load("local_search.npz")
    space = space_of(z)
    plan = neighbor_plan(space)
    inputs = space.get_input_parameters()
    objs = space.get_input_parameters_objects()
    np.random.seed(0)
    for _ in range(200):
        conf = space.get_random_configuration(use_priors=False)
        nb = get_neighbors(conf, space)
        rows = []
        for j, e in plan:
            name = inputs[j]
            vals = list(objs[name].get_discrete_values())
            n = len(vals)
            idx = vals.index(conf[name])
            if space.get_type(name) == "categorical" or n <= 5:
                pick = e + 1 if (j > 0 and e >= idx) else e
            else:
                w0 = 0 if idx < 2 else (n - 5 if (n - idx) <= 2 else idx - 2)
                pick = w0 + (e + 1 if (j > 0 and e >= idx - w0) else e)
            row = [conf[p] for p in inputs]
            row[j] = vals[pick]
            rows.append(row)
        assert rows == nb


# ---- local_search host logic against recorded reference runs (real / integer / mixed / discrete spaces) -------------
def _oracle_plugin():
    """An acquisition plug-in with the reference's exact values that runs without a GPU (oracle = checker): used to pin
    the HOST side of local_search -- RNG stream of the neighbourhoods, seed selection, duplicate filter, final pick."""
    from oracle import oracle
    from hypermapper_b200 import models as hm

    def f(configurations, **p):
        fas = p.setdefault("_oracle_forests", {o: oracle.forest_arrays_from_model(m)
                                               for o, m in p["regression_models"].items()})
        X = np.ascontiguousarray(hm.encode_host(configurations, p["param_space"]).T)
        vals = oracle.rf_acquisition(p["acquisition_function"], fas, X, p["objective_weights"],
                                     p["scalarization_method"], p["objective_limits"], p["iteration_number"],
                                     p["data_array"])
        return list(np.asarray(vals, dtype=np.float64)), [1] * len(configurations)
    return f


@pytest.mark.parametrize("tag", ["branin_s0", "branin_s1", "branin_s2", "mixed_s0", "mixed_s1", "mixed_s2",
                                 "discrete_s0", "discrete_s1", "discrete_s2", "discrete_s3"])
def test_local_search_host_logic_replays_reference_runs(tag):
    """Every configuration the reference's local_search evaluated, in the same order, the same values (the plug-in is
    the bit-exact oracle) and the same pick -- on spaces whose neighbourhoods consume numpy's RNG stream too."""
    import ls_replay
    from test_api_gpu import FrozenForest
    z = ls_replay.load()
    plugin = _oracle_plugin()

    def wrapped(configurations, **p):
        p.pop("_oracle_forests", None)
        return plugin(configurations, **p)
    out = ls_replay.replay(z, tag, wrapped, FrozenForest)
    ls_replay.check(*out, rel=0.0)


def test_local_search_goldens_discriminate():
    """the recorded runs do not all end in the same configuration (a climb that ignored seed / acquisition would fail)"""
    import ls_replay
    z = ls_replay.load()
    for name in ("branin", "mixed", "discrete"):
        ends = {tuple(z[t + "_best"].tolist()) for t in ls_replay.TAGS if t.startswith(name)}
        assert len(ends) >= 3, (name, ends)


def test_checked_build_reports_itself():
    """the normal build answers -1 to hm_checked_failures (no device-side checks compiled in); the symbol exists"""
    from hypermapper_b200 import _lib
    L = ctypes.CDLL(_lib.SO_PATH)
    L.hm_checked_failures.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    if "checked" not in os.path.basename(_lib.SO_PATH):
        assert L.hm_checked_failures(None, None) == -1
