"""Packed candidates (all-discrete spaces): one 64-bit code per candidate, the split test a shift + unsigned compare on
the code (csrc/hm_pack.cuh, hm_forest_attach_packing, hm_rf_eval_packed).  Every output must be bit-identical to the
fp32 path on the encoded form of the same candidates -- sklearn's decision (double)x32 <= threshold (models.py:174-182 ->
Tree._apply_dense) -- including thresholds that are not float32-representable, thresholds equal to an encoded value, and
splits that cannot separate any two values of their parameter."""
import json

import numpy as np
import pytest

from test_api_gpu import FrozenForest, load
from test_rf_gpu import _synthetic_forest

pytestmark = pytest.mark.gpu


def _space(params):
    from bench_support.space_lite import SpaceLite
    return SpaceLite({"optimization_objectives": ["v"], "input_parameters": params})


SPACES = {
    "c3": None,        # the bench's C3 scenario: 6 ordinals (4..64 values) + 4 categoricals (2, 3, 4, 8): 41 bits
    "odd": {
        "o1": {"parameter_type": "ordinal", "values": [7]},                                  # a single value
        "o3": {"parameter_type": "ordinal", "values": [0.5, 1.5, 2.5]},
        "c1": {"parameter_type": "categorical", "values": ["only"]},
        "o5": {"parameter_type": "ordinal", "values": [1, 2, 4, 8, 16]},
        "c5": {"parameter_type": "categorical", "values": list("abcde")},
        "o100": {"parameter_type": "ordinal", "values": list(range(100))},
        "c2": {"parameter_type": "categorical", "values": ["p", "q"]},
    },
    # 5 x 8 one-hot bits + 3 + 6 = 49 bits: split tests reach below bit 32 -> the 64-bit-shift ("wide") kernel variant
    "wide": dict([("k%d" % i, {"parameter_type": "categorical", "values": list("abcdefghi")}) for i in range(5)] +
                 [("o8", {"parameter_type": "ordinal", "values": [1, 2, 4, 8, 16, 32, 64, 128]}),
                  ("o33", {"parameter_type": "ordinal", "values": list(range(33))})]),
    # every ordinal a power of two, no categorical with three or more values: a FAST layout WITHOUT a sentinel field,
    # i.e. the predicated walk (kind 1) instead of the self-looping leaves (kind 3) that "c3" and "odd" get
    "pow2": {
        "a2": {"parameter_type": "ordinal", "values": [0.25, 0.75]},
        "b4": {"parameter_type": "ordinal", "values": [1, 2, 3, 4]},
        "c2": {"parameter_type": "categorical", "values": ["x", "y"]},
        "d16": {"parameter_type": "ordinal", "values": list(range(16))},
        "e8": {"parameter_type": "ordinal", "values": [1, 2, 4, 8, 16, 32, 64, 128]},
        "f2": {"parameter_type": "categorical", "values": ["u", "v"]},
    },
}
PACKED_KIND = {"c3": 3, "odd": 3, "wide": 2, "pow2": 1}


def _make(name):
    from bench_support import workloads
    from bench_support.space_lite import SpaceLite
    from hypermapper_b200.space_device import DeviceSpace
    space = SpaceLite(workloads.scenario_c3()) if name == "c3" else _space(SPACES[name])
    return space, DeviceSpace(space)


def _raw(space, n, rng):
    from bench_support import workloads
    cols = workloads.raw_uniform_columns(space, n, rng)
    return np.stack([np.asarray(cols[p], dtype=np.float64) for p in space.get_input_parameters()], axis=1)


def _edge_thresholds(f, x32_rows, rng):
    """move a third of the thresholds onto / next to encoded values and a few outside the value range"""
    thr = f["threshold"]
    feat = f["feature"]
    for i in np.nonzero(f["left"] >= 0)[0]:
        u = rng.random()
        vals = np.unique(x32_rows[:, feat[i]]).astype(np.float64)
        v = float(vals[rng.integers(0, len(vals))])
        if u < 0.12:
            thr[i] = v                                        # exactly an encoded value (x <= thr goes LEFT)
        elif u < 0.2:
            thr[i] = np.nextafter(v, -np.inf)                 # a hair below: not float32-representable
        elif u < 0.28:
            thr[i] = np.nextafter(v, np.inf)
        elif u < 0.31:
            thr[i] = -0.3                                     # every value goes right
        elif u < 0.34:
            thr[i] = 7.5                                      # every value goes left
    return f


@pytest.mark.parametrize("name,N,T,dlo,dhi", [("c3", 5000, 7, 3, 9), ("odd", 4000, 5, 2, 8), ("c3", 300_000, 40, 8, 12),
                                              ("odd", 140_000, 48, 7, 11), ("c3", 3000, 3, 13, 15),
                                              ("wide", 6000, 6, 3, 9), ("wide", 150_000, 40, 8, 11),
                                              ("wide", 2500, 3, 13, 15),
                                              ("pow2", 5000, 6, 3, 9), ("pow2", 140_000, 40, 8, 11),
                                              ("pow2", 2500, 3, 13, 15)])
def test_packed_path_is_bit_identical_to_fp32_path(name, N, T, dlo, dhi):
    """resident / streamed + coherence pre-pass / walked-in-L2 forests on two spaces, edge thresholds included"""
    import torch
    from hypermapper_b200 import forest, _lib
    space, ds = _make(name)
    assert ds.packed_bits == {"c3": 37, "odd": 1 + 2 + 0 + 3 + 4 + 7 + 1, "wide": 49, "pow2": 1 + 2 + 1 + 4 + 3 + 1}[name]
    rng = np.random.default_rng(N + T)
    raw = _raw(space, N, rng)
    x32, _ = ds.encode(raw)
    codes = ds.encode_packed(raw)
    assert np.array_equal(codes.cpu().numpy().view(np.uint64), ds.pack_codes_host(raw))
    assert np.array_equal(ds.unpack_codes_host(codes.cpu().numpy()), raw)
    rows32 = x32.T.cpu().numpy()
    fs = [_edge_thresholds(_synthetic_forest(rng, T, ds.n_encoded, dlo, dhi), rows32, rng) for _ in range(2)]
    dfs = [forest.DeviceForest(forest.tables_from_arrays(**f)) for f in fs]
    forest.attach_packing(dfs, ds)
    assert [f.packed_kind for f in dfs] == [PACKED_KIND[name]] * 2
    acq = _lib.make_acq_params("EI", "tchebyshev", [0.3, 0.7], [0.1, 0.2], 0.0)
    a = forest.rf_eval(dfs, x32, want_leaves=True, want_mean=True, want_var=True, want_pred=True)
    b = forest.rf_eval_packed(dfs, codes, want_leaves=True, want_mean=True, want_var=True, want_pred=True)
    for m in range(2):
        assert torch.equal(a["leaves"][m], b["leaves"][m])
    for k in ("mean", "var", "pred"):
        assert torch.equal(a[k], b[k]), k
    # the scoring launch (no leaf output: large pools take the coherence pre-pass), TS reads the plain-predict column
    for kind in ("EI", "TS"):
        p = _lib.make_acq_params(kind, "tchebyshev", [0.3, 0.7], [0.1, 0.2], 0.0)
        sa = forest.rf_eval(dfs, x32, acq=p)["score"]
        sb = forest.rf_eval_packed(dfs, codes, acq=p)["score"]
        assert torch.equal(sa.view(torch.int64), sb.view(torch.int64)), kind
        # ... and a launch that also returns leaves never takes the resident-forest mode that reads the leaf records from
        # shared memory (NVM = 2), so this pins that mode's scores as well wherever the small cases above select it
        sc = forest.rf_eval_packed(dfs, codes, want_leaves=True, acq=p)["score"]
        assert torch.equal(sb.view(torch.int64), sc.view(torch.int64)), kind


@pytest.mark.parametrize("name,N,T", [("c3", 200_000, 24), ("odd", 50_000, 12), ("pow2", 50_000, 12), ("wide", 50_000, 12)])
def test_malformed_codes_stay_inside_the_forest(name, N, T):
    """codes that are no candidate of the space (random 64-bit words: several one-hot bits set, ranks beyond the value
    list) must still end in SOME leaf of every tree -- in particular with self-looping leaves (kind 3), whose sentinel test
    is only guaranteed false for well-formed codes: the kernel replaces a code that breaks the sentinel bound by the
    all-zero code.  The checked build (tools/checked_run.py) verifies every staged-node address on the way."""
    import torch
    from hypermapper_b200 import forest
    space, ds = _make(name)
    rng = np.random.default_rng(T)
    raw = _raw(space, 4096, rng)
    x32, _ = ds.encode(raw)
    fs = [_synthetic_forest(rng, T, ds.n_encoded, 6, 11) for _ in range(2)]
    dfs = [forest.DeviceForest(forest.tables_from_arrays(**f)) for f in fs]
    forest.attach_packing(dfs, ds)
    garbage = torch.from_numpy(rng.integers(-2 ** 63, 2 ** 63, size=N, dtype=np.int64)).cuda()
    out = forest.rf_eval_packed(dfs, garbage, want_leaves=True, want_mean=True)
    torch.cuda.synchronize()
    for m in range(2):
        leaves = out["leaves"][m].cpu().numpy()
        for t in range(T):
            lo, hi = fs[m]["tree_off"][t], fs[m]["tree_off"][t + 1]
            is_leaf = fs[m]["left"][lo:hi] < 0
            assert np.all((leaves[t] >= 0) & (leaves[t] < hi - lo)) and np.all(is_leaf[leaves[t]])
    assert bool(torch.isfinite(out["mean"]).all())
    # and the big-pool launch (coherence pre-pass walks the key trees with the same codes)
    big = garbage.repeat(max(1, (1 << 18) // N + 1))
    forest.rf_eval_packed(dfs, big, want_mean=True)
    torch.cuda.synchronize()


def test_packed_path_on_the_reference_golden_forest():
    """the fitted forest and the pools of the reference's discrete local_search golden: packed == fp32 == oracle"""
    import torch
    from hypermapper_b200 import forest
    from hypermapper_b200.space_device import DeviceSpace
    from bench_support.space_lite import SpaceLite
    from oracle import oracle
    z = load("local_search.npz")
    space = SpaceLite(json.loads(str(z["config_json"])))
    ds = DeviceSpace(space)
    model = FrozenForest(z, "forest_Value_")
    df = forest.DeviceForest(forest.forest_tables(model))
    forest.attach_packing([df], ds)
    raw = np.concatenate([z["pool_uniform_EI_3"], z["pool_prior_EI_3"]], axis=0)
    out = forest.rf_eval_packed([df], ds.encode_packed(raw), want_leaves=True, want_mean=True, want_var=True)
    x32, _ = ds.encode(raw)
    ref = forest.rf_eval([df], x32, want_leaves=True, want_mean=True, want_var=True)
    assert torch.equal(out["leaves"][0], ref["leaves"][0]) and torch.equal(out["mean"], ref["mean"])
    assert torch.equal(out["var"], ref["var"])
    fa = oracle.forest_arrays_from_model(model)
    leaf = oracle.rf_apply(fa, x32.T.cpu().numpy().astype(np.float64))
    assert np.array_equal(out["leaves"][0].cpu().numpy().astype(np.int64), leaf)
    assert np.array_equal(out["mean"][0].cpu().numpy(), oracle.rf_prediction(fa, leaf))


def test_device_pool_as_packed_codes_is_the_same_pool():
    """hm_sample_pool_packed draws the very candidates hm_sample_pool draws (same Philox stream)"""
    space, ds = _make("c3")
    a = ds.sample_pool(1234, 50_000, True, want_raw=True)
    b = ds.sample_pool_packed(1234, 50_000, True, want_raw=True)
    assert np.array_equal(a["raw"].cpu().numpy(), b["raw"].cpu().numpy())
    assert np.array_equal(ds.unpack_codes_host(b["codes"].cpu().numpy()), a["raw"].T.cpu().numpy())
    assert np.array_equal(ds.encode_packed(a["raw"].T.contiguous()).cpu().numpy(), b["codes"].cpu().numpy())


def test_packed_host_entry_and_errors():
    """hm_rf_score_packed_host (8 bytes per candidate over PCIe) == the device path, across chunk boundaries; spaces
    with real parameters refuse to pack; a forest that was not attached refuses packed calls"""
    import torch
    from hypermapper_b200 import forest, _lib
    from hypermapper_b200.scoring import HostScorer
    from hypermapper_b200.space_device import DeviceSpace
    space, ds = _make("c3")
    rng = np.random.default_rng(8)
    raw = _raw(space, 9000, rng)
    fs = [_synthetic_forest(rng, 12, ds.n_encoded, 3, 8) for _ in range(2)]
    dfs = [forest.DeviceForest(forest.tables_from_arrays(**f)) for f in fs]
    codes_h = ds.pack_codes_host(raw)
    acq = _lib.make_acq_params("EI", "tchebyshev", [0.3, 0.7], [0.1, 0.2], 0.0)
    with pytest.raises(_lib.HmError):
        forest.rf_eval_packed(dfs, torch.from_numpy(codes_h.view(np.int64)).cuda(), acq=acq)
    forest.attach_packing(dfs, ds)
    want = forest.rf_eval_packed(dfs, torch.from_numpy(codes_h.view(np.int64)).cuda(), acq=acq)["score"]
    wv, wi = forest.topk(want, 20)
    scores = np.empty(len(raw))
    _, tv, ti = HostScorer(ds.n_encoded, chunk=2048).score_packed(dfs, codes_h, acq, k=20, scores_out=scores)
    assert np.array_equal(scores, want.cpu().numpy())
    assert np.array_equal(ti, wi.cpu().numpy()) and np.array_equal(tv, wv.cpu().numpy())
    mixed = DeviceSpace(_space({"r": {"parameter_type": "real", "values": [0, 1]},
                                "o": {"parameter_type": "ordinal", "values": [1, 2, 3]}}))
    assert mixed.packed_bits == 0
    with pytest.raises(_lib.HmError):
        mixed.encode_packed(np.array([[0.5, 2.0]]))
    with pytest.raises(ValueError):
        ds.encode_packed(np.full((3, ds.D), 12345.0))


def test_device_pools_of_discrete_spaces_are_packed_and_score_identically():
    """local_search's device pools on an all-discrete space are 64-bit codes; acquisition_scores_device walks them with
    the packed kernel and returns the very scores the fp32 encoded pool gets; the encoded matrix is still available on
    demand (feasibility classifier / GP consumers)."""
    import torch
    from hypermapper_b200 import random_scalarizations as rs
    from bench_support import workloads
    from hypermapper_b200.models import EncodedCandidates
    from hypermapper_b200.space_device import DevicePool, device_space_for
    wl = workloads.RFWorkload(dict(workloads.scenario_c3(), models={"model": "random_forest", "number_of_trees": 12}),
                              n_train=300, seed=3)
    ds = device_space_for(wl.space)
    pool = DevicePool(ds, 99, False, 200_000)
    assert pool.candidates.codes is not None and pool.candidates.x32 is None
    args = ("EI", None, wl.objective_weights, wl.models, wl.space, "tchebyshev", wl.objective_limits, 5, wl.data_array,
            "random_forest")
    got = rs.acquisition_scores_device(args[0], pool.candidates, *args[2:])
    ref_pool = ds.sample_pool(99, 200_000, False)
    want = rs.acquisition_scores_device(args[0], EncodedCandidates(x32=ref_pool["x32"]), *args[2:])
    assert torch.equal(got.view(torch.int64), want.view(torch.int64))
    assert torch.equal(pool.candidates.get("float32"), ref_pool["x32"])
