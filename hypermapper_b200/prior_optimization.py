"""Prior-guided acquisition (BOPrO), same module-level names and signatures as the reference's prior_optimization.py:

    compute_probability_from_prior      prior_optimization.py:65-95      K7 hm_prior_prob
    estimate_prior_limits               prior_optimization.py:98-123
    compute_probability_from_model      prior_optimization.py:126-158
    compute_EI_from_posteriors          prior_optimization.py:161-374    hm_bopro_ratio / hm_minmax / hm_bopro_finish
    prior_guided_optimization           prior_optimization.py:377-476

The per-candidate arithmetic runs on the device; the host keeps the O(1) state the reference keeps (the running
normalisation limits, which the reference updates IN PLACE in the lists the caller passes, :173-178, :318-327).
"""
import ctypes
import sys

import numpy as np

from . import _lib
from . import models
from .encoding import raw_matrix
from .space_device import device_space_for


def _raw_matrix(configurations, param_space):
    return np.array(raw_matrix(configurations, param_space), dtype=np.float64, copy=True)


def prior_probability_device(configurations, param_space, objective_weights):
    """compute_probability_from_prior, result left on the GPU (CUDA fp64 tensor [N])."""
    objectives = param_space.get_optimization_parameters()
    w = [objective_weights[o] for o in objectives]
    if isinstance(configurations, models.EncodedCandidates):
        if configurations.raw_cm is None:
            raise _lib.HmError("prior-weighted acquisitions need the raw parameter values of the candidates")
        return device_space_for(param_space).prior_prob(configurations.raw_cm, w, col_major=True)
    return device_space_for(param_space).prior_prob(_raw_matrix(configurations, param_space), w)


def compute_probability_from_prior(configurations, param_space, objective_weights):
    """prior_optimization.py:65-95 -> list of probabilities."""
    if param_space.get_estimate_prior_flags()[0]:
        raise _lib.HmError("estimated (multivariate KDE) priors are evaluated by the reference's space.pdf; "
                           "they have no device form")
    return prior_probability_device(configurations, param_space, objective_weights).cpu().tolist()


def estimate_prior_limits(param_space, prior_limit_estimation_points, objective_weights):
    """prior_optimization.py:98-123"""
    uniform = param_space.random_sample_configurations_without_repetitions({}, prior_limit_estimation_points,
                                                                           use_priors=False)
    prior = param_space.random_sample_configurations_without_repetitions({}, prior_limit_estimation_points,
                                                                         use_priors=True)
    p = prior_probability_device(uniform + prior, param_space, objective_weights)
    lo, hi = _minmax(p, ignore_inf=False)
    return [lo, hi]


def _minmax(t, ignore_inf):
    torch = _lib.require_cuda()
    out = torch.empty((2,), dtype=torch.float64, device=t.device)
    ws = torch.empty((2,), dtype=torch.int64, device=t.device)
    rc = _lib.lib().hm_minmax(_lib.dptr(t), t.numel(), 1 if ignore_inf else 0, _lib.dptr(out), _lib.dptr(ws),
                              _lib.stream_ptr())
    _lib.check(rc, "hm_minmax")
    lo, hi = out.cpu().tolist()
    return lo, hi


def compute_probability_from_model(model_means, model_stds, param_space, objective_weights, threshold,
                                   compute_bad=True):
    """prior_optimization.py:126-158 (host arrays in, host array out; O(N M) glue kept for API completeness --
    compute_EI_from_posteriors does this on the device inside hm_bopro_ratio)."""
    from scipy import stats
    objs = param_space.get_optimization_parameters()
    probabilities = np.ones(len(model_means[objs[0]]))
    for o in objs:
        c = stats.norm.cdf((threshold[o] - model_means[o]) / model_stds[o])
        probabilities *= ((1 - c) if compute_bad else c) ** objective_weights[o]
    return probabilities


def posterior_scores_device(configurations, param_space, objective_weights, objective_limits, threshold,
                            iteration_number, model_weight, regression_models, classification_model, model_type,
                            good_prior_normalization_limits, posterior_floor=10 ** -8,
                            posterior_normalization_limits=None, debug=False):
    """compute_EI_from_posteriors with the results left on the GPU -> (scores [N], feasibility indicator [N] | None).
    `configurations` may be a list of dicts or an [N][D] raw array."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    objectives = list(param_space.get_optimization_parameters())
    ds = device_space_for(param_space)
    if isinstance(configurations, models.EncodedCandidates):
        # a device pool: generated inside the parameter ranges, the clipping of :189-197 is the identity
        enc = configurations
        N = len(enc)
        prior = prior_probability_device(enc, param_space, objective_weights)
    else:
        raw = _raw_matrix(configurations, param_space)
        # :189-197 clip every configuration to the parameter range
        objs = param_space.get_input_parameters_objects()
        for j, p in enumerate(param_space.get_input_parameters()):
            raw[:, j] = np.maximum(np.minimum(raw[:, j], objs[p].get_max()), objs[p].get_min())
        raw_d = ds.raw_to_device(raw)
        N = raw_d.shape[0]
        prior = ds.prior_prob(raw_d, [objective_weights[o] for o in objectives])
        enc = models.EncodedCandidates(*ds.encode(raw_d, want32=True, want64=(model_type == "gaussian_process")))

    bp = _lib.BoproParams()
    bp.n_objectives = len(objectives)
    for m, o in enumerate(objectives):
        bp.weight[m] = float(objective_weights[o])
        bp.threshold[m] = float(threshold[o])
    bp.prior_mode = 0
    if good_prior_normalization_limits is not None and N > 0:
        lo, hi = _minmax(prior, ignore_inf=False)
        good_prior_normalization_limits[0] = min(good_prior_normalization_limits[0], lo)
        good_prior_normalization_limits[1] = max(good_prior_normalization_limits[1], hi)
        if good_prior_normalization_limits[0] == good_prior_normalization_limits[1]:
            bp.prior_mode = 2
        else:
            bp.prior_mode = 1
            bp.prior_lo, bp.prior_hi = good_prior_normalization_limits[0], good_prior_normalization_limits[1]
    bp.posterior_floor = float(posterior_floor)
    discrete = all(param_space.get_type(p) != "real" for p in param_space.get_input_parameters())
    bp.discrete_space = 1 if discrete else 0
    bp.discrete_divisor = float(param_space.get_discrete_space_size() - 1) if discrete else 1.0
    bp.model_exponent = iteration_number / model_weight

    mean, var = models.mean_and_variance_device(enc, regression_models, model_type, param_space)
    feas_prob = None
    if classification_model is not None:
        feas_prob = models.feasibility_probability_device(enc.get("float32"), classification_model, param_space)
    ratio = torch.empty((N,), dtype=torch.float64, device="cuda")
    rc = L.hm_bopro_ratio(ctypes.byref(bp), _lib.dptr(prior), _lib.dptr(mean), _lib.dptr(var), N, N, _lib.dptr(ratio),
                          _lib.stream_ptr())
    _lib.check(rc, "hm_bopro_ratio")
    mode, lo, hi = 0, 0.0, 1.0
    if posterior_normalization_limits is not None and N > 0:
        rlo, rhi = _minmax(ratio, ignore_inf=True)
        posterior_normalization_limits[0] = min(posterior_normalization_limits[0], rlo)
        posterior_normalization_limits[1] = max(posterior_normalization_limits[1], rhi)
        if posterior_normalization_limits[0] == posterior_normalization_limits[1]:
            mode = 2
        else:
            mode, lo, hi = 1, posterior_normalization_limits[0], posterior_normalization_limits[1]
    out = torch.empty((N,), dtype=torch.float64, device="cuda")
    feas_out = torch.empty((N,), dtype=torch.float64, device="cuda") if feas_prob is not None else None
    rc = L.hm_bopro_finish(_lib.dptr(ratio), N, mode, lo, hi, float(posterior_floor), _lib.dptr(feas_prob),
                           _lib.dptr(out), _lib.dptr(feas_out), _lib.stream_ptr())
    _lib.check(rc, "hm_bopro_finish")
    return out, feas_out


def compute_EI_from_posteriors(configurations, param_space, objective_weights, objective_limits, threshold,
                               iteration_number, model_weight, regression_models, classification_model, model_type,
                               good_prior_normalization_limits, posterior_floor=10 ** -8,
                               posterior_normalization_limits=None, debug=False):
    """prior_optimization.py:161-374 -> (list of values, feasibility indicator).  Like the reference it clips the
    passed configurations to the parameter ranges in place and updates the two limit lists in place."""
    objs = param_space.get_input_parameters_objects()
    if isinstance(configurations, list) and len(configurations) > 0 and isinstance(configurations[0], dict):
        for p in param_space.get_input_parameters():
            lo, hi = objs[p].get_min(), objs[p].get_max()
            for c in configurations:
                c[p] = max(min(c[p], hi), lo)
    out, feas = posterior_scores_device(configurations, param_space, objective_weights, objective_limits, threshold,
                                        iteration_number, model_weight, regression_models, classification_model,
                                        model_type, good_prior_normalization_limits, posterior_floor,
                                        posterior_normalization_limits, debug)
    values = out.cpu().tolist()
    feasibility = feas.cpu().numpy() if feas is not None else [1] * len(values)
    return values, feasibility


def _device_scores(candidates, **p):
    p = {k: v for k, v in p.items() if k != "number_of_cpus"}
    return posterior_scores_device(candidates, **p)[0]


compute_EI_from_posteriors.device_scores = _device_scores


def prior_guided_optimization(config, data_array, param_space, fast_addressing_of_data_array, regression_models,
                              iteration_number, objective_weights, objective_limits, classification_model=None,
                              profiling=None, acquisition_function_optimizer="local_search"):
    """prior_optimization.py:377-476: pack the parameters and run the acquisition optimiser."""
    from .local_search import local_search
    fp = {
        "param_space": param_space,
        "iteration_number": iteration_number,
        "regression_models": regression_models,
        "classification_model": classification_model,
        "objective_weights": objective_weights,
        "objective_limits": objective_limits,
        "model_type": config["models"]["model"],
        "model_weight": config["model_posterior_weight"],
        "posterior_floor": config["posterior_computation_lower_limit"],
        "threshold": {o: np.quantile(data_array[o], config["model_good_quantile"])
                      for o in param_space.get_optimization_parameters()},
    }
    if param_space.get_prior_normalization_flag() is True:
        fp["good_prior_normalization_limits"] = estimate_prior_limits(
            param_space, config["prior_limit_estimation_points"], objective_weights)
    else:
        fp["good_prior_normalization_limits"] = None
    if classification_model is not None:
        fp["posterior_normalization_limits"] = [float("inf"), float("-inf")]
    if acquisition_function_optimizer == "local_search":
        _, best = local_search(config["local_search_starting_points"], config["local_search_random_points"],
                               param_space, fast_addressing_of_data_array, False, compute_EI_from_posteriors, fp,
                               config["scalarization_key"], config["number_of_cpus"], previous_points=data_array,
                               profiling=profiling)
        return best
    if acquisition_function_optimizer == "cma_es":
        raise _lib.HmError("cma_es is out of scope of the batched hot path (SURVEY.md section 2); use the reference's "
                           "cma_es with hypermapper_b200.prior_optimization.compute_EI_from_posteriors as its objective")
    print("Unrecognized acquisition function optimizer:", acquisition_function_optimizer)
    raise SystemExit
