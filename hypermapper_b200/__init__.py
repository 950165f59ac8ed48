"""hypermapper_b200: B200-native (sm_100a) surrogate prediction + acquisition scoring hot path of HyperMapper.

Python host code mirrors the reference's module-level API for that path (models / random_scalarizations /
local_search / compute_pareto); the arithmetic runs in hand-written CUDA kernels behind the C-ABI declared in
include/hm_b200.h.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import HmError, build  # noqa: F401

__all__ = ["HmError", "build"]


def install(hypermapper_package=None):
    """Drop this package in underneath an importable reference `hypermapper` package: rebinds the hot-path entry points
    of hypermapper.models / random_scalarizations / local_search / compute_pareto / pibo / prior_optimization (and the
    RFModel methods) to the
    B200 implementations.  Call it after `import hypermapper...` and before `optimizer.optimize(...)`; the scenario JSON
    and schema.json are untouched (backend selection must not add JSON keys: schema.json:441 additionalProperties=false).
    Returns the list of rebound names."""
    import importlib

    from . import compute_pareto as b_cp
    from . import local_search as b_ls
    from . import models as b_models
    from . import pibo as b_pibo
    from . import prior_optimization as b_po
    from . import random_scalarizations as b_rs

    def mod(name):
        if hypermapper_package is not None:
            return getattr(hypermapper_package, name)
        return importlib.import_module("hypermapper." + name)

    done = []

    def bind(target, name, fn):
        setattr(target, name, fn)
        done.append("%s.%s" % (getattr(target, "__name__", target), name))

    r_models = mod("models")
    for name in ("compute_model_mean_and_uncertainty", "compute_rf_prediction_mean_and_uncertainty",
                 "compute_gp_prediction_mean_and_uncertainty", "model_prediction", "model_probabilities"):
        bind(r_models, name, getattr(b_models, name))
    bind(r_models.RFModel, "get_leaves_per_sample", b_models.rf_get_leaves_per_sample)
    bind(r_models.RFModel, "compute_rf_prediction", b_models.rf_compute_rf_prediction)
    bind(r_models.RFModel, "compute_rf_prediction_variance", b_models.rf_compute_rf_prediction_variance)
    r_rs = mod("random_scalarizations")
    for name in ("run_acquisition_function", "EI", "ucb", "thompson_sampling", "random_scalarizations"):
        bind(r_rs, name, getattr(b_rs, name))
    r_ls = mod("local_search")
    bind(r_ls, "local_search", b_ls.local_search)
    bind(r_ls, "get_neighbors", b_ls.get_neighbors)
    bind(r_models, "local_search", b_ls.local_search)          # models.minimize_posterior_mean imports it by name
    r_cp = mod("compute_pareto")
    for name in ("sequential_is_pareto_efficient", "sequential_is_pareto_efficient_dumb", "parallel_is_pareto_efficient",
                 "compute_pareto"):
        bind(r_cp, name, getattr(b_cp, name))
    # prior-guided acquisitions (pibo.py, prior_optimization.py); both modules import the optimiser and each other's
    # helpers by name at import time
    r_po = mod("prior_optimization")
    for name in ("compute_probability_from_prior", "estimate_prior_limits", "compute_probability_from_model",
                 "compute_EI_from_posteriors", "prior_guided_optimization"):
        bind(r_po, name, getattr(b_po, name))
    bind(r_po, "local_search", b_ls.local_search)
    r_pibo = mod("pibo")
    for name in ("prior_weighted_acquisition_function", "run_pibo"):
        bind(r_pibo, name, getattr(b_pibo, name))
    bind(r_pibo, "compute_probability_from_prior", b_po.compute_probability_from_prior)
    bind(r_pibo, "run_acquisition_function", b_rs.run_acquisition_function)
    bind(r_pibo, "local_search", b_ls.local_search)
    # bo.py binds its three optimisation methods at import time (bo.py:15-17) -> rebind there too if it is loaded
    import sys
    bo = sys.modules.get("hypermapper.bo")
    if bo is not None:
        for name, fn in (("random_scalarizations", b_rs.random_scalarizations), ("run_pibo", b_pibo.run_pibo),
                         ("prior_guided_optimization", b_po.prior_guided_optimization)):
            if hasattr(bo, name):
                bind(bo, name, fn)
    return done
