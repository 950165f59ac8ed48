"""CPU: the oracle restatement against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Bit-exact unless stated."""
import json
import os

import numpy as np
import pytest

from oracle import oracle

from conftest import GOLDEN


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def forests_of(z):
    objectives = json.loads(str(z["objectives_json"]))
    return objectives, {o: oracle.ForestArrays.from_npz(z, "forest_%s_" % o) for o in objectives}


def columns_of(z):
    params = json.loads(str(z["params_json"]))
    cand = z["candidates"]
    cols = {}
    for j, p in enumerate(params):
        c = cand[:, j]
        if p["type"] in ("integer", "categorical"):
            cols[p["name"]] = [int(v) for v in c]
        elif p["type"] == "ordinal":
            ints = all(float(v).is_integer() for v in p["values"])
            cols[p["name"]] = [int(v) if ints else float(v) for v in c]
        else:
            cols[p["name"]] = [float(v) for v in c]
    return params, cols


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
def test_encoding_matches_reference(name):
    z = load(name)
    params, cols = columns_of(z)
    enc, _ = oracle.encode(cols, params)
    assert np.array_equal(enc, z["encoded"])


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
def test_rf_leaves_mean_var_bit_exact(name):
    z = load(name)
    objectives, forests = forests_of(z)
    X = z["encoded"]
    for o in objectives:
        fa = forests[o]
        leaf = oracle.rf_apply(fa, X)
        assert np.array_equal(leaf, z["leaves_" + o])
        mean = oracle.rf_prediction(fa, leaf)
        assert np.array_equal(mean, z["mean_" + o])
        var = oracle.rf_prediction_variance(fa, leaf, mean)
        assert np.array_equal(var, z["var_" + o])
        assert np.array_equal(np.sqrt(var), z["std_" + o])
        assert np.array_equal(oracle.rf_sklearn_predict(fa, leaf), z["sk_predict_" + o])
        m2, v2, p2 = oracle.rf_mean_var_omp(fa, X, want_pred=True)
        assert np.array_equal(m2, mean) and np.array_equal(v2, var) and np.array_equal(p2, z["sk_predict_" + o])


def _acq_inputs(z):
    objectives, forests = forests_of(z)
    data = {o: z["data_" + o].tolist() for o in objectives}
    weights = {o: float(w) for o, w in zip(objectives, z["weights"])}
    limits = {o: [float(a), float(b)] for o, (a, b) in zip(objectives, z["stored_limits"])}
    return objectives, forests, data, weights, limits


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
@pytest.mark.parametrize("scal", ["linear", "tchebyshev", "modified_tchebyshev"])
def test_data_array_scalarization(name, scal):
    z = load(name)
    objectives, _, data, weights, limits = _acq_inputs(z)
    s, lim = oracle.compute_data_array_scalarization(data, weights, limits, scal)
    assert np.array_equal(s, z["data_scalarization_" + scal])
    assert np.array_equal(np.array([lim[o] for o in objectives]), z["data_scalarization_limits_" + scal])


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
@pytest.mark.parametrize("acq", ["EI", "UCB", "TS"])
@pytest.mark.parametrize("scal", ["linear", "tchebyshev", "modified_tchebyshev"])
def test_acquisition_bit_exact(name, acq, scal):
    z = load(name)
    objectives, forests, data, weights, limits = _acq_inputs(z)
    vals = oracle.rf_acquisition(acq, forests, z["encoded"], weights, scal, limits, int(z["iteration_number"]), data)
    assert np.array_equal(vals, z["acq_%s_%s" % (acq, scal)])


@pytest.mark.parametrize("name", ["rf_mixed.npz", "rf_branin.npz"])
@pytest.mark.parametrize("k", [1, 20, 64])
def test_get_min_configurations(name, k):
    z = load(name)
    vals = z["acq_EI_tchebyshev"]
    idx = oracle.get_min_indices(vals, k)
    assert np.array_equal(idx, z["topk_idx_%d" % k])
    assert np.array_equal(vals[idx], z["topk_val_%d" % k])
    assert np.array_equal(oracle.get_min_indices_fast(vals, k), idx)


def test_get_min_indices_ties_nan_negzero():
    rng = np.random.default_rng(0)
    v = np.round(rng.standard_normal(5000), 1)
    v[[17, 400]] = np.nan
    v[[3, 9]] = -0.0
    v[[5]] = 0.0
    for k in (1, 7, 100):
        assert np.array_equal(oracle.get_min_indices(v, k), oracle.get_min_indices_fast(v, k))


PARETO_CASES = ["uniform3", "uniform2", "uniform4", "ties3", "ties2", "anticorr2", "dups3", "nan3", "single", "quirk3",
                "quirk3_swapped", "blackscholes"]


@pytest.mark.parametrize("case", PARETO_CASES)
def test_pareto_matches_reference(case):
    z = load("pareto.npz")
    costs = z["costs_" + case]
    mask = oracle.pareto(costs)
    assert np.array_equal(mask, z["mask_" + case])
    if costs.shape[0] <= 4000:
        with np.errstate(all="ignore"):
            assert np.array_equal(oracle.pareto_numpy(costs), z["mask_" + case])


def test_pareto_known_answers_of_reference_tests():
    """tests/data/BlackScholes_real_pareto.csv (88 rows): survivors of the oracle == the reference's golden CSV."""
    z = load("pareto.npz")
    costs = z["costs_blackscholes"]
    got = np.array(sorted(costs[oracle.pareto(costs)].tolist()))
    assert np.array_equal(got, z["kat_blackscholes"])


def _branin2(x1, x2):
    a, b, c, r, s, t = 1.0, 5.1 / (4.0 * np.pi * np.pi), 5.0 / np.pi, 6.0, 10.0, 1.0 / (8.0 * np.pi)
    y = a * (x2 - b * x1 * x1 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s
    return y


def test_pareto_pass1_is_superset_and_order_independent():
    z = load("pareto.npz")
    costs = z["costs_ties3"]
    p1 = oracle.pareto_pass1(costs)
    assert np.all(p1[z["mask_ties3"]])
    perm = np.random.default_rng(1).permutation(len(costs))
    p1p = oracle.pareto_pass1(costs[perm])
    assert np.array_equal(p1p, p1[perm])


def test_gp_restatement_self_consistency():
    """GP oracle is parity-unpinned (GPy absent); check algebraic identities of the restatement."""
    rng = np.random.default_rng(2)
    X = rng.random((40, 3))
    y = np.sin(3 * X[:, 0]) + X[:, 1] ** 2 - X[:, 2]
    st = oracle.gp_fit(X, y, [0.4, 0.7, 0.9], 1.3, 1e-4)
    mean, var = oracle.gp_predict(st, X)
    # interpolates the training data up to the noise level, variance ~ noise at training points
    assert np.max(np.abs(mean - y)) < 5e-2
    assert np.all(var >= 1e-11) and np.max(var) < 1e-2 * np.var(y) + 1e-3
    Xs = rng.random((200, 3))
    m2, v2 = oracle.gp_predict(st, Xs)
    # direct formula with explicit inverse
    K = oracle.matern52(st.X, st.X, st.lengthscale, st.variance) + (st.noise + 1e-8) * np.eye(40)
    Kx = oracle.matern52(st.X, Xs, st.lengthscale, st.variance)
    Ki = np.linalg.inv(K)
    yn = (y - y.mean()) / y.std()
    mu = Kx.T @ Ki @ yn * y.std() + y.mean()
    vv = (st.variance - np.einsum("ij,ik,kj->j", Kx, Ki, Kx) + st.noise) * y.std() ** 2
    assert np.allclose(m2, mu, rtol=1e-6, atol=1e-8)
    assert np.allclose(v2, np.maximum(vv, 1e-11), rtol=1e-4, atol=1e-9)


# ---- GP oracle pinned to an independent implementation --------------------------------------------------------------
GP_TAGS = ["c2_noise", "c2_nonoise", "c5gp", "small_rbf", "sf2"]


@pytest.mark.parametrize("tag", GP_TAGS)
def test_gp_oracle_agrees_with_sklearn_gpr(tag):
    """oracle.gp_predict (restatement of GPy's exact inference, models.py:441-459, 997-1024) against scikit-learn's
    GaussianProcessRegressor on the same fixed hyper-parameters (tests/golden/gp_sklearn.npz, made by make_golden.py
    gp_case): C2 shape (n = 256, D = 6, noise 1e-4 and 0), C5-GP shape (n = 512, D = 32), RBF, sf2 != 1.
    sklearn builds the kernel from explicit pairwise distances and has its own Cholesky / normaliser code, so this is
    an implementation the oracle's author did not write.  The one known GPy detail mirrored on the sklearn side is the
    1e-8 jitter GPy adds to the noise variance (sklearn: alpha = 1e-8).
    Tolerances: mean 1e-9 relative wherever |mean| > 1e-3 std(y) (absolute 1e-12 std(y) below that: the mean is a
    cancelling sum); variance 1e-6 relative -- the worst case is a candidate ON a training point with zero noise, where
    var = sf2 - |L^-1 k|^2 cancels down to the 1e-8 jitter."""
    z = np.load(os.path.join(GOLDEN, "gp_sklearn.npz"), allow_pickle=False)
    X, y, ls = z[tag + "_X"], z[tag + "_y"], z[tag + "_lengthscale"]
    sf2, noise = z[tag + "_hyper"]
    st = oracle.gp_fit(X, y, ls, float(sf2), float(noise), kernel=str(z[tag + "_kernel"]))
    m, v = oracle.gp_predict(st, z[tag + "_candidates"])
    sm, sv = z[tag + "_sk_mean"], z[tag + "_sk_var"]
    big = np.abs(sm) > 1e-3 * y.std()
    assert np.all(np.abs(m - sm)[big] <= 1e-9 * np.abs(sm)[big])
    assert np.all(np.abs(m - sm)[~big] <= 1e-12 * y.std())
    assert np.all(np.abs(v - sv) <= 1e-6 * sv)
    assert np.all(np.abs(v - sv)[10:] <= 1e-9 * sv[10:])          # away from the training points: 1e-9


@pytest.mark.parametrize("kind", ["branin_grid", "ordinal_branin"])
def test_pareto_oracle_reference_known_answer_grids(kind):
    """the two grid KATs of the reference's tests/test_system.py:36-64 (98 / 12 rows)"""
    import pareto_kat
    z = np.load(os.path.join(GOLDEN, "pareto.npz"), allow_pickle=False)
    pareto_kat.check(z, kind, oracle.pareto)
