"""GP surrogate state -> device handle (hm_gp_create) and the predict launch (hm_gp_mean_var).

Fitting (hyper-parameter optimisation) is out of scope and stays with GPy / the reference
(models.py:441-457); this module only reads the fitted state:
    X, ARD lengthscales, kernel variance sf2, Gaussian noise variance, the lower Cholesky factor L of
    K + (noise + 1e-8) I, alpha = K_y^-1 y_normalised, and the Standardize normaliser (mean, std).
`GPState.from_gpy(model)` extracts them from a GPy GPRegression; `GPState.from_hyperparameters` builds the same
state for fixed hyper-parameters (synthetic workloads / tests, where GPy is not installed).
"""
import ctypes
import weakref

import numpy as np

from . import _lib


class GPState:
    def __init__(self, X, lengthscale, variance, noise, L, alpha, y_mean, y_std, kernel="matern52"):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.lengthscale = np.ascontiguousarray(np.broadcast_to(np.asarray(lengthscale, dtype=np.float64),
                                                                (self.X.shape[1],)))
        self.variance = float(variance)
        self.noise = float(noise)
        self.L = np.ascontiguousarray(L, dtype=np.float64)
        self.alpha = np.ascontiguousarray(alpha, dtype=np.float64).reshape(-1)
        self.y_mean = float(y_mean)
        self.y_std = float(y_std)
        self.kernel = kernel

    @staticmethod
    def from_gpy(model):
        """GPy.models.GPRegression(X, y, Matern52(d, ARD=True), normalizer=True) as built at models.py:444-450."""
        kern = model.kern
        name = type(kern).__name__.lower()
        kernel = "matern52" if "matern52" in name else ("rbf" if "rbf" in name else None)
        if kernel is None:
            raise _lib.HmError("unsupported GPy kernel %s" % type(kern).__name__)
        post = model.posterior
        norm = getattr(model, "normalizer", None)
        y_mean = float(np.asarray(norm.mean).reshape(-1)[0]) if norm is not None else 0.0
        y_std = float(np.asarray(norm.std).reshape(-1)[0]) if norm is not None else 1.0
        return GPState(np.asarray(model.X), np.asarray(kern.lengthscale), float(np.asarray(kern.variance).reshape(-1)[0]),
                       float(np.asarray(model.likelihood.variance).reshape(-1)[0]), np.asarray(post.woodbury_chol),
                       np.asarray(post.woodbury_vector), y_mean, y_std, kernel)

    @staticmethod
    def from_hyperparameters(X, y, lengthscale, variance, noise, kernel="matern52"):
        """Exact-inference state for FIXED hyper-parameters: Standardize normaliser, K_y = K + (noise + 1e-8) I,
        L = chol(K_y), alpha = K_y^-1 y_norm."""
        from scipy import linalg
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        ls = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (X.shape[1],))
        y_mean, y_std = float(y.mean()), float(y.std())
        yn = (y - y_mean) / y_std
        Xs = X / ls
        sq = np.sum(np.square(Xs), 1)
        r2 = -2.0 * Xs.dot(Xs.T) + (sq[:, None] + sq[None, :])
        np.fill_diagonal(r2, 0.0)
        r = np.sqrt(np.clip(r2, 0, np.inf))
        if kernel == "matern52":
            K = variance * (1 + np.sqrt(5.0) * r + 5.0 / 3 * r ** 2) * np.exp(-np.sqrt(5.0) * r)
        else:
            K = variance * np.exp(-0.5 * r ** 2)
        L = np.linalg.cholesky(K + (noise + 1e-8) * np.eye(len(y)))
        alpha = linalg.cho_solve((L, True), yn)
        return GPState(X, ls, variance, noise, L, alpha, y_mean, y_std, kernel)


class DeviceGP:
    def __init__(self, state):
        from scipy import linalg
        _lib.require_cuda()
        n, D = state.X.shape
        Xs = np.ascontiguousarray(state.X / state.lengthscale)                     # GPy: X / lengthscale
        Linv = np.ascontiguousarray(linalg.solve_triangular(state.L, np.eye(n), lower=True))
        h = ctypes.c_void_p()
        rc = _lib.lib().hm_gp_create(n, D, _lib.hptr(Xs), _lib.hptr(state.lengthscale), _lib.hptr(state.alpha),
                                     _lib.hptr(Linv), state.variance, state.noise, state.y_mean, state.y_std,
                                     0 if state.kernel == "matern52" else 1, ctypes.byref(h))
        _lib.check(rc, "hm_gp_create")
        self.handle = h
        self.n, self.D = n, D

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().hm_gp_destroy(h)
            except Exception:
                pass
            self.handle = None


_cache = weakref.WeakKeyDictionary()


def device_gp(model):
    """DeviceGP for a GPState, a DeviceGP, or a fitted GPy model (cached on the object)."""
    if isinstance(model, DeviceGP):
        return model
    try:
        hit = _cache.get(model)
    except TypeError:
        hit = None
    if hit is not None:
        return hit
    state = model if isinstance(model, GPState) else GPState.from_gpy(model)
    dg = DeviceGP(state)
    try:
        _cache[model] = dg
    except TypeError:
        pass
    return dg


def gp_mean_var(dgp, X64, mean_out, var_out):
    """X64: CUDA fp64 [D][N] column-major raw encoded candidates; writes mean_out/var_out (CUDA fp64 [N])."""
    torch = _lib.require_cuda()
    assert X64.dtype == torch.float64 and X64.is_cuda and X64.is_contiguous()
    D, N = X64.shape
    if D != dgp.D:
        raise ValueError("candidates have %d encoded features, the GP was fitted on %d" % (D, dgp.D))
    if N == 0:
        return
    rc = _lib.lib().hm_gp_mean_var(dgp.handle, _lib.dptr(X64), N, N, _lib.dptr(mean_out), _lib.dptr(var_out),
                                   _lib.stream_ptr())
    _lib.check(rc, "hm_gp_mean_var")
