"""GPU parity for K2 (GP posterior mean / variance + EI) against the numpy/scipy oracle restatement of GPy's exact
inference AND against scikit-learn's GaussianProcessRegressor (tests/golden/gp_sklearn.npz: an independent exact-inference
implementation the oracle is pinned to, tests/test_oracle.py).  The real GPy is not installable here (oracle/__init__.py).
Tolerance: 1e-5 relative (north_star) on mean, variance and acquisition values, per element:
  mean     |d| <= 1e-5 * max(|mean|, 1e-3 * std(y))   -- mu = Kx.alpha is a cancelling sum; where |mean| < 1e-3 std(y)
           the bound is absolute (1e-8 std(y)), the accuracy any summation order of the n terms can promise
  variance |d| <= 1e-5 * var + 8 n eps sf2 std(y)^2    -- var = sf2 - |L^-1 k|^2 is a cancellation: its absolute rounding
           error is ~n eps sf2 whatever the implementation (the oracle's own BLAS result moves by that much between
           summation orders); the floor only matters for candidates on top of training points with ~zero noise"""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-5


def _setup(n, D, noise, seed, kernel="matern52"):
    from hypermapper_b200 import gp
    from bench_support import workloads
    from oracle import oracle
    rng = np.random.default_rng(seed)
    X = rng.random((n, D))
    y = workloads.hartmann6(X) if D == 6 else np.sin(3 * X[:, 0]) + (X ** 2).sum(1)
    ls = np.random.default_rng(seed + 1).uniform(0.3, 1.0, D)
    st_o = oracle.gp_fit(X, y, ls, 1.0, noise, kernel=kernel)
    st = gp.GPState.from_hyperparameters(X, y, ls, 1.0, noise, kernel=kernel)
    return st_o, st


def assert_mean_var(m, v, mean_ref, var_ref, n, st):
    """per-element 1e-5 relative (see the module docstring for the two floors)"""
    assert np.all(np.abs(m - mean_ref) <= REL * np.maximum(np.abs(mean_ref), 1e-3 * st.y_std))
    floor = 8 * n * np.finfo(np.float64).eps * st.variance * st.y_std ** 2
    assert np.all(np.abs(v - var_ref) <= REL * var_ref + floor)


@pytest.mark.parametrize("tag", ["c2_noise", "c2_nonoise", "c5gp", "small_rbf", "sf2"])
def test_gp_mean_var_vs_sklearn_golden(tag):
    """K2 against scikit-learn's GaussianProcessRegressor outputs (C2 and C5-GP shapes among them)"""
    import torch
    from conftest import GOLDEN
    from hypermapper_b200 import gp
    z = np.load(os.path.join(GOLDEN, "gp_sklearn.npz"), allow_pickle=False)
    X, y, ls = z[tag + "_X"], z[tag + "_y"], z[tag + "_lengthscale"]
    sf2, noise = z[tag + "_hyper"]
    st = gp.GPState.from_hyperparameters(X, y, ls, float(sf2), float(noise), kernel=str(z[tag + "_kernel"]))
    Xs = z[tag + "_candidates"]
    N = len(Xs)
    X64 = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda()
    mean = torch.empty(N, dtype=torch.float64, device="cuda")
    var = torch.empty_like(mean)
    gp.gp_mean_var(gp.DeviceGP(st), X64, mean, var)
    assert_mean_var(mean.cpu().numpy(), var.cpu().numpy(), z[tag + "_sk_mean"], z[tag + "_sk_var"], len(X), st)


def rel_err(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.mark.parametrize("n,D,noise,N,kernel", [(256, 6, 1e-4, 20000, "matern52"), (256, 6, 0.0, 20000, "matern52"),
                                                (200, 6, 1e-4, 5001, "matern52"), (17, 3, 1e-2, 1000, "matern52"),
                                                (512, 32, 1e-4, 4000, "matern52"), (700, 10, 1e-3, 2000, "matern52"),
                                                (300, 4, 1e-4, 3000, "rbf"), (3, 2, 1e-2, 100, "matern52")])
def test_gp_mean_var_vs_oracle(n, D, noise, N, kernel):
    import torch
    from hypermapper_b200 import gp
    from oracle import oracle
    st_o, st = _setup(n, D, noise, seed=n + D, kernel=kernel)
    Xs = np.random.default_rng(4).random((N, D))
    Xs[:min(5, n)] = st.X[:5]                      # candidates ON training points: the ill-conditioned end of the variance
    mean_o, var_o = oracle.gp_predict(st_o, Xs)
    X64 = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda()
    mean = torch.empty(N, dtype=torch.float64, device="cuda")
    var = torch.empty_like(mean)
    gp.gp_mean_var(gp.device_gp(st), X64, mean, var)
    m, v = mean.cpu().numpy(), var.cpu().numpy()
    assert_mean_var(m, v, mean_o, var_o, n, st)
    assert np.all(v >= 1e-11)


def test_gp_ei_matches_oracle_c2_shape():
    """C2: Hartmann-6, n = 256, EI single objective, through run_acquisition_function-level plumbing."""
    import torch
    from hypermapper_b200 import gp, models, _lib
    from oracle import oracle
    st_o, st = _setup(256, 6, 1e-4, seed=2)
    N = 50000
    Xs = np.random.default_rng(9).random((N, 6))
    mean_o, var_o = oracle.gp_predict(st_o, Xs)
    y = (st_o.alpha * 0).tolist()
    data = {"Value": (np.random.default_rng(1).standard_normal(50)).tolist()}
    w = {"Value": 1.0}
    lim = {"Value": [min(data["Value"]), max(data["Value"])]}
    want = oracle.EI({"Value": mean_o}, {"Value": var_o}, data, w, "tchebyshev", lim)
    X64 = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda()
    mean = torch.empty((1, N), dtype=torch.float64, device="cuda")
    var = torch.empty_like(mean)
    gp.gp_mean_var(gp.device_gp(st), X64, mean[0], var[0])
    _, lim2 = oracle.compute_data_array_scalarization(data, w, lim, "tchebyshev")
    p = _lib.make_acq_params("EI", "tchebyshev", [1.0], [oracle.ei_f_min(data, "Value", lim2)], 0.0)
    got = models.acq_score_device(p, mean, var).cpu().numpy()
    assert rel_err(got, want) <= REL


def test_gp_host_scoring_entry_matches_device_path():
    """hm_gp_score_host (host candidates -> chunked H2D -> hm_gp_mean_var -> hm_acq_score -> D2H) equals the
    device-resident path bit for bit, top-k included, across chunk boundaries."""
    import torch
    from hypermapper_b200 import gp, models, forest, _lib
    from hypermapper_b200.scoring import HostScorer
    rng = np.random.default_rng(5)
    n, D, N = 96, 5, 7000
    X = rng.random((n, D))
    y = np.sin(X.sum(1) * 3) + 0.1 * rng.standard_normal(n)
    state = gp.GPState.from_hyperparameters(X, y, rng.uniform(0.3, 1.0, D), 1.3, 1e-4, "matern52")
    dgp = gp.DeviceGP(state)
    Xc = np.ascontiguousarray(rng.random((D, N)))
    acq = _lib.make_acq_params("EI", "tchebyshev", [1.0], [0.4], 0.0)
    xd = torch.from_numpy(Xc).cuda()
    mean = torch.empty((1, N), dtype=torch.float64, device="cuda")
    var = torch.empty_like(mean)
    gp.gp_mean_var(dgp, xd, mean[0], var[0])
    want = models.acq_score_device(acq, mean, var)
    wv, wi = forest.topk(want, 20)
    scorer = HostScorer(D, chunk=2048)
    scores = np.empty(N)
    _, tv, ti = scorer.score_gp([dgp], Xc, acq, k=20, scores_out=scores)
    assert np.array_equal(scores, want.cpu().numpy())
    assert np.array_equal(ti, wi.cpu().numpy()) and np.array_equal(tv, wv.cpu().numpy())


def test_gp_full_size_properties_c2():
    """BASELINE C2 size (Hartmann-6, n = 256, 2^20 candidates): size-independent properties.
    (1) a candidate's mean / variance do not depend on its position in the pool or on its tile mates (a pool built from
        a 2^16 block repeated 16 times under different shuffles gives every copy bit-identical results);
    (2) the stable top-k of the EI scores is the block's best candidate, lowest index first;
    (3) a sample agrees with the oracle within 1e-5."""
    import torch
    from hypermapper_b200 import gp, models, forest, _lib
    from oracle import oracle
    st_o, st = _setup(256, 6, 1e-4, seed=2)
    dgp = gp.device_gp(st)
    rng = np.random.default_rng(4)
    base_n, reps = 1 << 16, 16
    base = torch.from_numpy(np.ascontiguousarray(rng.random((6, base_n)))).cuda()
    perms = [torch.randperm(base_n, device="cuda") for _ in range(reps)]
    pool = torch.cat([base[:, p] for p in perms], dim=1).contiguous()
    N = pool.shape[1]
    mean = torch.empty((1, N), dtype=torch.float64, device="cuda")
    var = torch.empty_like(mean)
    gp.gp_mean_var(dgp, pool, mean[0], var[0])
    bm = torch.empty((1, base_n), dtype=torch.float64, device="cuda")
    bv = torch.empty_like(bm)
    gp.gp_mean_var(dgp, base, bm[0], bv[0])
    for r, p in enumerate(perms):
        assert torch.equal(mean[0, r * base_n:(r + 1) * base_n], bm[0, p]), r
        assert torch.equal(var[0, r * base_n:(r + 1) * base_n], bv[0, p]), r
    acq = _lib.make_acq_params("EI", "tchebyshev", [1.0], [0.6], 0.0)
    score = models.acq_score_device(acq, mean, var)
    bscore = models.acq_score_device(acq, bm, bv)
    tv, ti = forest.topk(score, 16)
    best = bscore.min()
    where = torch.nonzero(score == best).view(-1)[:16]
    assert torch.equal(ti[: len(where)], where) and bool((tv[: len(where)] == best).all())
    sub = rng.choice(base_n, 2000, replace=False)
    mo, vo = oracle.gp_predict(st_o, base[:, torch.from_numpy(sub).cuda()].T.cpu().numpy())
    m, v = bm[0].cpu().numpy()[sub], bv[0].cpu().numpy()[sub]
    assert_mean_var(m, v, mo, vo, 256, st)


@pytest.mark.parametrize("n,D,noise,N", [(1000, 8, 1e-3, 3000), (1537, 20, 1e-4, 700), (2048, 6, 1e-2, 9000)])
def test_gp_large_models_beyond_the_fused_kernel(n, D, noise, N):
    """n > 768 training points (a 1000-evaluation BO run gets there): the chunked Kx + triangular DMMA GEMM path,
    against the oracle (pinned to scikit-learn); chunk boundary (N > 8192) and padding (n not a multiple of 64) included"""
    import torch
    from hypermapper_b200 import gp
    from oracle import oracle
    st_o, st = _setup(n, D, noise, seed=n + D)
    Xs = np.random.default_rng(4).random((N, D))
    Xs[:5] = st.X[:5]
    Xs[7, 0] = np.nan
    mean_o, var_o = oracle.gp_predict(st_o, np.where(np.isnan(Xs), 0.5, Xs))
    X64 = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda()
    mean = torch.empty(N, dtype=torch.float64, device="cuda")
    var = torch.empty_like(mean)
    gp.gp_mean_var(gp.DeviceGP(st), X64, mean, var)
    m, v = mean.cpu().numpy(), var.cpu().numpy()
    assert np.isnan(m[7]) and np.isnan(v[7])                # NaN coordinates: GPy's K is NaN there
    ok = np.arange(N) != 7
    assert_mean_var(m[ok], v[ok], mean_o[ok], var_o[ok], n, st)


def test_gp_large_path_agrees_with_fused_kernel(monkeypatch):
    """the two K2 paths on the same model (HM_B200_GP_FORCE_BIG routes a small model through the large-n path)"""
    import torch
    from hypermapper_b200 import gp
    _, st = _setup(300, 6, 1e-4, seed=5)
    Xs = np.random.default_rng(6).random((5000, 6))
    X64 = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda()
    outs = []
    for force in (False, True):
        if force:
            monkeypatch.setenv("HM_B200_GP_FORCE_BIG", "1")
        mean = torch.empty(5000, dtype=torch.float64, device="cuda")
        var = torch.empty_like(mean)
        gp.gp_mean_var(gp.DeviceGP(st), X64, mean, var)
        outs.append((mean.cpu().numpy(), var.cpu().numpy()))
    assert_mean_var(outs[1][0], outs[1][1], outs[0][0], outs[0][1], 300, st)


def test_gp_c5_shape_quarter_million_candidates_vs_oracle():
    """BASELINE configs[4], GP leg: the C5-GP shape (n = 512 training points, D_enc = 32) on 2^18 candidates against
    the oracle (chunked on the host), mean / variance per element within 1e-5, EI top-20 identical."""
    import torch
    from hypermapper_b200 import forest, gp, models, _lib
    from oracle import oracle
    n, D, N = 512, 32, 1 << 18
    rng = np.random.default_rng(12)
    X = rng.random((n, D))
    y = np.sin(X[:, :8].sum(1)) + (X[:, 8:] ** 2).sum(1) * 0.1
    ls = np.random.default_rng(3).uniform(0.3, 1.0, D) * 3.0
    st_o = oracle.gp_fit(X, y, ls, 1.0, 1e-4)
    st = gp.GPState.from_hyperparameters(X, y, ls, 1.0, 1e-4)
    Xs = np.random.default_rng(5).random((N, D))
    mean_o, var_o = np.empty(N), np.empty(N)
    for lo in range(0, N, 1 << 15):
        mean_o[lo:lo + (1 << 15)], var_o[lo:lo + (1 << 15)] = oracle.gp_predict(st_o, Xs[lo:lo + (1 << 15)])
    X64 = torch.from_numpy(np.ascontiguousarray(Xs.T)).cuda()
    mean = torch.empty((1, N), dtype=torch.float64, device="cuda")
    var = torch.empty_like(mean)
    gp.gp_mean_var(gp.DeviceGP(st), X64, mean[0], var[0])
    assert_mean_var(mean[0].cpu().numpy(), var[0].cpu().numpy(), mean_o, var_o, n, st)
    data = {"Value": [float(v) for v in y]}
    w = {"Value": 1.0}
    lim = {"Value": [float(y.min()), float(y.max())]}
    want = oracle.EI({"Value": mean_o}, {"Value": var_o}, data, w, "tchebyshev", lim)
    _, lim2 = oracle.compute_data_array_scalarization(data, w, lim, "tchebyshev")
    p = _lib.make_acq_params("EI", "tchebyshev", [1.0], [oracle.ei_f_min(data, "Value", lim2)], 0.0)
    got = models.acq_score_device(p, mean, var)
    assert rel_err(got.cpu().numpy(), want) <= REL
    assert np.array_equal(forest.topk(got, 20)[1].cpu().numpy(), oracle.get_min_indices_fast(want, 20))
