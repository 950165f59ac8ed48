"""GPU parity: the CUDA forest / acquisition / top-k path (through the C-ABI) against the reference's golden
vectors and against the CPU oracle on seeded inputs."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

REL_VAR = 1e-10   # variance: mean**2 is libm pow(x,2) in the reference, x*x on the device (<= 1 ulp apart)
REL_ACQ = 1e-5    # north_star tolerance for acquisition values


def _torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def device_forests(z):
    from hypermapper_b200 import forest
    objectives = json.loads(str(z["objectives_json"]))
    out = []
    for o in objectives:
        pre = "forest_%s_" % o
        tabs = forest.tables_from_arrays(z[pre + "tree_off"], z[pre + "left"], z[pre + "right"], z[pre + "feature"],
                                         z[pre + "threshold"], z[pre + "leaf_mean"], z[pre + "leaf_var"],
                                         z[pre + "value"], int(z[pre + "n_features"]))
        out.append(forest.DeviceForest(tabs))
    return objectives, out


def xcm(z):
    torch = _torch()
    enc32 = np.ascontiguousarray(z["encoded"].astype(np.float32).T)
    return torch.from_numpy(enc32).cuda()


def assert_rel(a, b, rel):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape
    both_nan = np.isnan(a) & np.isnan(b)
    with np.errstate(all="ignore"):
        err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    err[both_nan] = 0
    err[(a == b)] = 0
    assert np.all(err <= rel), "max rel err %g" % np.max(err)


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
def test_leaves_mean_var_vs_reference_golden(name):
    from hypermapper_b200 import forest
    z = load(name)
    objectives, forests = device_forests(z)
    out = forest.rf_eval(forests, xcm(z), want_leaves=True, want_mean=True, want_var=True, want_pred=True)
    for m, o in enumerate(objectives):
        assert np.array_equal(out["leaves"][m].cpu().numpy().astype(np.int64), z["leaves_" + o])   # bit-exact
        assert np.array_equal(out["mean"][m].cpu().numpy(), z["mean_" + o])                       # bit-exact
        assert_rel(out["var"][m].cpu().numpy(), z["var_" + o], REL_VAR)
        assert np.array_equal(out["pred"][m].cpu().numpy(), z["sk_predict_" + o])                 # bit-exact


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
@pytest.mark.parametrize("acq", ["EI", "UCB", "TS"])
@pytest.mark.parametrize("scal", ["linear", "tchebyshev", "modified_tchebyshev"])
def test_fused_acquisition_vs_reference_golden(name, acq, scal):
    from hypermapper_b200 import forest, _lib
    from oracle import oracle
    z = load(name)
    objectives, forests = device_forests(z)
    data = {o: z["data_" + o].tolist() for o in objectives}
    weights = {o: float(w) for o, w in zip(objectives, z["weights"])}
    limits = {o: [float(a), float(b)] for o, (a, b) in zip(objectives, z["stored_limits"])}
    _, lim = oracle.compute_data_array_scalarization(data, weights, limits, scal)
    w = oracle.reciprocate_weights(weights) if scal == "modified_tchebyshev" else weights
    fmin = [oracle.ei_f_min(data, o, lim) for o in objectives]
    it = int(z["iteration_number"])
    beta = float(np.sqrt(0.125 * np.log(2 * it + 1)))
    p = _lib.make_acq_params(acq, scal, [w[o] for o in objectives], fmin, beta)
    out = forest.rf_eval(forests, xcm(z), acq=p)
    assert_rel(out["score"].cpu().numpy(), z["acq_%s_%s" % (acq, scal)], REL_ACQ)


@pytest.mark.parametrize("k", [1, 20, 64])
def test_topk_matches_get_min_configurations_golden(k):
    from hypermapper_b200 import forest
    torch = _torch()
    z = load("rf_mixed.npz")
    vals = torch.from_numpy(z["acq_EI_tchebyshev"]).cuda()
    ov, oi = forest.topk(vals, k)
    assert np.array_equal(oi.cpu().numpy(), z["topk_idx_%d" % k])
    assert np.array_equal(ov.cpu().numpy(), z["topk_val_%d" % k])


@pytest.mark.parametrize("n,k", [(1, 1), (5, 20), (1000, 20), (100000, 20), (3_000_000, 64), (3_000_000, 1024)])
def test_topk_vs_oracle_ties_nan(n, k):
    from hypermapper_b200 import forest
    from oracle import oracle
    torch = _torch()
    rng = np.random.default_rng(n + k)
    v = np.round(rng.standard_normal(n), 2)      # heavy ties
    if n > 100:
        v[rng.integers(0, n, 3)] = np.nan
        v[rng.integers(0, n, 5)] = -0.0
        v[rng.integers(0, n, 2)] = -np.inf
    ov, oi = forest.topk(torch.from_numpy(v).cuda(), k, index_base=7)
    want = oracle.get_min_indices_fast(v, k)
    got = oi.cpu().numpy()
    assert np.array_equal(got[: len(want)] - 7, want)
    assert np.all(got[len(want):] == -1)
    gv = ov.cpu().numpy()[: len(want)]
    assert np.array_equal(np.isnan(gv), np.isnan(v[want]))
    assert np.array_equal(gv[~np.isnan(gv)], v[want][~np.isnan(gv)] + 0.0)


def test_topk_descending_input_worst_case():
    from hypermapper_b200 import forest
    torch = _torch()
    n = 1_000_000
    v = torch.arange(n, 0, -1, dtype=torch.float64, device="cuda")
    ov, oi = forest.topk(v, 20)
    assert oi.cpu().tolist() == list(range(n - 1, n - 21, -1))


def test_topk_merge_equals_global_topk():
    from hypermapper_b200 import forest
    torch = _torch()
    rng = np.random.default_rng(5)
    v = np.round(rng.standard_normal(400_000), 3)
    vals = torch.from_numpy(v).cuda()
    k = 20
    gv, gi = forest.topk(vals, k)
    parts_v, parts_i = [], []
    for r in range(4):     # four "ranks", contiguous shards, global indices
        lo, hi = r * 100_000, (r + 1) * 100_000
        pv, pi = forest.topk(vals[lo:hi].contiguous(), k, index_base=lo)
        parts_v.append(pv)
        parts_i.append(pi)
    mv, mi = forest.topk_merge(torch.cat(parts_v), torch.cat(parts_i), k)
    assert torch.equal(mi, gi) and torch.equal(mv, gv)


def _synthetic_forest(rng, T, D, depth_lo, depth_hi):
    """random binary trees in sklearn layout with random Hutter tables"""
    offs, left, right, feat, thr, lm, lv, val = [0], [], [], [], [], [], [], []
    for _ in range(T):
        l, r, f, th = [], [], [], []

        def grow(depth):
            i = len(l)
            l.append(-1)
            r.append(-1)
            f.append(-2)
            th.append(-2.0)
            if depth < depth_lo or (depth < depth_hi and rng.random() < 0.7):
                f[i] = int(rng.integers(0, D))
                th[i] = float(rng.random())
                l[i] = grow(depth + 1)
                r[i] = grow(depth + 1)
            return i

        grow(0)
        n = len(l)
        offs.append(offs[-1] + n)
        left += l
        right += r
        feat += f
        thr += th
        lm += list(rng.standard_normal(n))
        lv += list(rng.random(n) * (rng.random(n) > 0.2))
        val += list(rng.standard_normal(n))
    return dict(tree_off=np.array(offs), left=np.array(left), right=np.array(right), feature=np.array(feat),
                threshold=np.array(thr), leaf_mean=np.array(lm), leaf_var=np.array(lv), value=np.array(val), n_features=D)


@pytest.mark.parametrize("T,D,dlo,dhi,N", [(3, 2, 1, 3, 5000), (64, 23, 4, 11, 20000), (300, 32, 6, 12, 6000),
                                           (17, 60, 2, 9, 3001), (2, 150, 1, 5, 777), (4, 5, 13, 15, 4000)])
def test_forest_kernel_vs_oracle_synthetic(T, D, dlo, dhi, N):
    """streamed (multi-group), resident, narrow-block and wide-feature configurations against the C oracle"""
    from hypermapper_b200 import forest
    from oracle import oracle
    torch = _torch()
    rng = np.random.default_rng(T * 1000 + D)
    fs = [_synthetic_forest(rng, T, D, dlo, dhi) for _ in range(2)]
    X = rng.random((N, D))
    X[:, 0] = np.round(X[:, 0], 1)
    fas = [oracle.ForestArrays(f["tree_off"], f["left"], f["right"], f["feature"], f["threshold"], f["leaf_mean"],
                               f["leaf_var"], f["value"], D) for f in fs]
    dfs = [forest.DeviceForest(forest.tables_from_arrays(**f)) for f in fs]
    Xd = torch.from_numpy(np.ascontiguousarray(X.astype(np.float32).T)).cuda()
    out = forest.rf_eval(dfs, Xd, want_leaves=True, want_mean=True, want_var=True, want_pred=True)
    for m, fa in enumerate(fas):
        leaf = oracle.rf_apply(fa, X)
        assert np.array_equal(out["leaves"][m].cpu().numpy().astype(np.int64), leaf)
        mean = oracle.rf_prediction(fa, leaf)
        assert np.array_equal(out["mean"][m].cpu().numpy(), mean)
        assert_rel(out["var"][m].cpu().numpy(), oracle.rf_prediction_variance(fa, leaf, mean), 1e-9)
        assert np.array_equal(out["pred"][m].cpu().numpy(), oracle.rf_sklearn_predict(fa, leaf))


@pytest.mark.parametrize("T,D,N", [(40, 23, 300_000), (24, 32, 200_001)])
def test_coherence_prepass_is_invisible(T, D, N):
    """hm_rf_eval buckets large pools by the leaves of two key trees and walks them through that permutation
    (hm_forest.cu, coherence pre-pass).  Every output must be bit-identical with the pre-pass off, and equal the
    C oracle on a sample."""
    from hypermapper_b200 import forest, _lib
    from oracle import oracle
    torch = _torch()
    rng = np.random.default_rng(T + D)
    fs = [_synthetic_forest(rng, T, D, 8, 12) for _ in range(2)]
    X = rng.random((N, D))
    X[:, 1] = np.round(X[:, 1], 1)
    dfs = [forest.DeviceForest(forest.tables_from_arrays(**f)) for f in fs]
    Xd = torch.from_numpy(np.ascontiguousarray(X.astype(np.float32).T)).cuda()
    acq = _lib.make_acq_params("EI", "tchebyshev", [0.3, 0.7], [0.1, 0.2], 0.0)
    outs = []
    try:
        for mode in (0, 1):
            for f in dfs:
                _lib.check(_lib.lib().hm_forest_set_coherence(f.handle, mode), "hm_forest_set_coherence")
            o = forest.rf_eval(dfs, Xd, want_mean=True, want_var=True, want_pred=True, acq=acq)
            outs.append({k: [t.cpu().numpy() for t in v] if isinstance(v, (list, tuple)) else v.cpu().numpy()
                         for k, v in o.items() if v is not None})
    finally:
        for f in dfs:
            _lib.lib().hm_forest_set_coherence(f.handle, -1)
    for k in outs[0]:
        a, b = outs[0][k], outs[1][k]
        if isinstance(a, list):
            assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b)), k
        else:
            assert np.array_equal(a, b, equal_nan=True), k
    sub = rng.choice(N, 4000, replace=False)
    for m, f in enumerate(fs):
        fa = oracle.ForestArrays(f["tree_off"], f["left"], f["right"], f["feature"], f["threshold"], f["leaf_mean"],
                                 f["leaf_var"], f["value"], D)
        leaf = oracle.rf_apply(fa, X[sub].astype(np.float32).astype(np.float64))
        assert np.array_equal(outs[1]["mean"][m][sub], oracle.rf_prediction(fa, leaf))


def test_threshold_rounding_edge_cases():
    """x32 <= thr32 must equal (double)x32 <= thr64 for thresholds that are not float32-representable."""
    from hypermapper_b200 import forest
    from oracle import oracle
    torch = _torch()
    xs = np.array([0.1, 0.3, 1.0, 1e-30, 16777217.0, -0.7], dtype=np.float32)
    thr, X = [], []
    for x in xs:
        d = float(x)
        for t in (d, np.nextafter(d, np.inf), np.nextafter(d, -np.inf), float(np.nextafter(x, np.float32(np.inf))),
                  float(np.nextafter(x, np.float32(-np.inf)))):
            thr.append(t)
            X.append(x)
    T = len(thr)
    # every tree tests every candidate, so each (x, threshold) pairing above is covered on the diagonal and more
    f = dict(tree_off=np.arange(0, 3 * T + 1, 3), left=np.tile([1, -1, -1], T), right=np.tile([2, -1, -1], T),
             feature=np.tile([0, -2, -2], T), threshold=np.repeat(thr, 3), leaf_mean=np.zeros(3 * T),
             leaf_var=np.zeros(3 * T), value=np.zeros(3 * T), n_features=1)
    Xn = np.array(X, dtype=np.float64).reshape(-1, 1)
    fa = oracle.ForestArrays(f["tree_off"], f["left"], f["right"], f["feature"], f["threshold"], f["leaf_mean"],
                             f["leaf_var"], f["value"], 1)
    df = forest.DeviceForest(forest.tables_from_arrays(**f))
    Xd = torch.from_numpy(np.ascontiguousarray(Xn.astype(np.float32).T)).cuda()
    out = forest.rf_eval([df], Xd, want_leaves=True)
    assert np.array_equal(out["leaves"][0].cpu().numpy().astype(np.int64), oracle.rf_apply(fa, Xn))


def test_forest_edge_cases_empty_single_leaf_mixed_layouts_eight_objectives():
    """Edge cases in one go: N = 0 and N = 1 pools; trees that are a single leaf; eight objectives in one launch; a
    call that mixes a shared-memory forest with one whose trees only fit in global memory (the two node layouts)."""
    from hypermapper_b200 import forest, _lib
    from oracle import oracle
    torch = _torch()
    rng = np.random.default_rng(77)
    D = 6
    small = [_synthetic_forest(rng, 5, D, 0, 4) for _ in range(7)]          # depth_lo = 0: some trees are one leaf
    assert any(np.any(np.diff(f["tree_off"]) == 1) for f in small)
    deep = _synthetic_forest(rng, 3, D, 14, 15)                               # > 16 k nodes per tree: not staged
    fs = small + [deep]
    dfs = [forest.DeviceForest(forest.tables_from_arrays(**f)) for f in fs]
    assert [int(_lib.lib().hm_forest_staged(d.handle)) for d in dfs] == [1] * 7 + [0]
    fas = [oracle.ForestArrays(f["tree_off"], f["left"], f["right"], f["feature"], f["threshold"], f["leaf_mean"],
                               f["leaf_var"], f["value"], D) for f in fs]
    acq = _lib.make_acq_params("UCB", "linear", [1.0 / 8] * 8, None, 0.7)
    for N in (0, 1, 2500):
        X = rng.random((N, D))
        Xd = torch.from_numpy(np.ascontiguousarray(X.astype(np.float32).T)).cuda().reshape(D, N)
        out = forest.rf_eval(dfs, Xd, want_leaves=True, want_mean=True, want_var=True, acq=acq)
        assert out["score"].shape == (N,) and out["mean"].shape == (8, N)
        for m, fa in enumerate(fas):
            leaf = oracle.rf_apply(fa, X.astype(np.float32).astype(np.float64)) if N else np.zeros((len(fa.tree_off) - 1, 0), dtype=np.int64)
            assert np.array_equal(out["leaves"][m].cpu().numpy().astype(np.int64), leaf)
            if N:
                assert np.array_equal(out["mean"][m].cpu().numpy(), oracle.rf_prediction(fa, leaf))
        if N:
            mean = out["mean"].cpu().numpy()
            var = out["var"].cpu().numpy()
            want = (mean / 8).sum(0) * 0          # UCB linear: sum w*mu - beta*sqrt(sum w*var)
            s = np.zeros(N)
            b = np.zeros(N)
            for m in range(8):
                s = s + (1.0 / 8) * mean[m]
                b = b + (1.0 / 8) * var[m]
            want = s - 0.7 * np.sqrt(b)
            assert_rel(out["score"].cpu().numpy(), want, 1e-12)


def test_full_size_pool_properties():
    """BASELINE size (2^24 candidates, D_enc = 23, two streamed forests): size-independent properties.
    (1) the score of a candidate does not depend on where it sits in the pool (a pool built from a 2^20 base block
        repeated 16 times under different shuffles scores every copy identically -- exercises the coherence pre-pass and
        the scatter back through its permutation at full size);
    (2) the stable top-k of that pool is the base block's best candidate, all 16 copies of it, lowest index first;
    (3) a sample agrees with the C oracle."""
    from hypermapper_b200 import forest, _lib
    from oracle import oracle
    torch = _torch()
    rng = np.random.default_rng(2024)
    D, T, base_n, reps = 23, 48, 1 << 20, 16
    fs = [_synthetic_forest(rng, T, D, 7, 11) for _ in range(2)]
    dfs = [forest.DeviceForest(forest.tables_from_arrays(**f)) for f in fs]
    assert all(int(_lib.lib().hm_forest_n_groups(d.handle)) > 1 for d in dfs)          # streamed, pre-pass active
    base = torch.from_numpy(np.round(rng.random((D, base_n)) * 16) / 16).to(torch.float32).cuda()
    perms = [torch.randperm(base_n, device="cuda") for _ in range(reps)]
    pool = torch.cat([base[:, p] for p in perms], dim=1).contiguous()                  # [D][2^24]
    acq = _lib.make_acq_params("EI", "tchebyshev", [0.4, 0.6], [0.3, 0.1], 0.0)
    score = forest.rf_eval(dfs, pool, acq=acq)["score"]
    ref = forest.rf_eval(dfs, base.contiguous(), acq=acq)["score"]
    for r, p in enumerate(perms):
        assert torch.equal(score[r * base_n:(r + 1) * base_n], ref[p]), r
    k = 16
    tv, ti = forest.topk(score, k)
    best = ref.min()
    n_best = int((ref == best).sum())
    where = torch.nonzero(score == best).view(-1)[:k]
    assert torch.equal(ti, where) and bool((tv == best).all()) and n_best * reps >= k
    sub = rng.choice(base_n, 3000, replace=False)
    Xs = base[:, torch.from_numpy(sub).cuda()].T.cpu().numpy().astype(np.float64)
    out = forest.rf_eval(dfs, base.contiguous(), want_mean=True)
    for m, f in enumerate(fs):
        fa = oracle.ForestArrays(f["tree_off"], f["left"], f["right"], f["feature"], f["threshold"], f["leaf_mean"],
                                 f["leaf_var"], f["value"], D)
        assert np.array_equal(out["mean"][m].cpu().numpy()[sub], oracle.rf_prediction(fa, oracle.rf_apply(fa, Xs)))


def test_resident_forests_with_leaf_records_in_shared_memory():
    """Forests that stay resident in shared memory (<= 2 groups over the whole call) carry their leaf records in their
    groups and are scored by a kernel mode that reads them from there (hm_forest.cu, NVM = 2).  That mode only runs
    when neither leaves nor the plain prediction are requested, so its means / variances / scores are compared bit for
    bit with the same call made WITH leaves (global-memory records) and with the C oracle: one single-group forest,
    two single-group forests (the second lands in the second buffer), one two-group forest, and a three-forest call
    that is no longer resident (falls back to the gathers)."""
    from hypermapper_b200 import forest, _lib
    from oracle import oracle
    torch = _torch()
    L = _lib.lib()
    rng = np.random.default_rng(2024)
    D, N = 3, 40_000
    X = rng.random((N, D))
    Xd = torch.from_numpy(np.ascontiguousarray(X.astype(np.float32).T)).cuda()
    small = [_synthetic_forest(rng, 10, D, 1, 5) for _ in range(3)]
    two_group = None
    # ~220 leaves per tree: 3.5 KB of nodes + 7 KB of leaf records against a node buffer of 107 KB (D = 3)
    for T in (12, 14, 16, 18, 10, 20, 8, 24):                   # the first size whose groups (records included) are exactly two
        cand = _synthetic_forest(rng, T, D, 7, 8)
        df = forest.DeviceForest(forest.tables_from_arrays(**cand))
        if int(L.hm_forest_n_groups(df.handle)) == 2 and int(L.hm_forest_leaf_resident(df.handle, 0)) == 1:
            two_group = cand
            break
    assert two_group is not None, "no two-group resident forest among the sizes tried"
    seen_groups = set()
    for fs, resident in (([small[0]], True), ([small[0], small[1]], True), ([two_group], True), (small, False)):
        dfs = [forest.DeviceForest(forest.tables_from_arrays(**f)) for f in fs]
        groups = sum(int(L.hm_forest_n_groups(d.handle)) for d in dfs)
        assert (groups <= 2) == resident and all(int(L.hm_forest_leaf_resident(d.handle, 0)) == 1 for d in dfs)
        seen_groups.add(groups)
        M = len(fs)
        acq = _lib.make_acq_params("EI", "tchebyshev", [1.0 / M] * M, [0.3] * M, 0.0)
        fast = forest.rf_eval(dfs, Xd, want_mean=True, want_var=True, acq=acq)
        slow = forest.rf_eval(dfs, Xd, want_leaves=True, want_mean=True, want_var=True, acq=acq)
        for k in ("mean", "var", "score"):
            assert np.array_equal(fast[k].cpu().numpy(), slow[k].cpu().numpy(), equal_nan=True), (k, groups)
        for m, f in enumerate(fs):
            fa = oracle.ForestArrays(f["tree_off"], f["left"], f["right"], f["feature"], f["threshold"], f["leaf_mean"],
                                     f["leaf_var"], f["value"], D)
            leaf = oracle.rf_apply(fa, X.astype(np.float32).astype(np.float64))
            assert np.array_equal(fast["mean"][m].cpu().numpy(), oracle.rf_prediction(fa, leaf))
    assert {1, 2} <= seen_groups
