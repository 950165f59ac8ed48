"""Acquisition functions of the hot path, same module-level names and signatures as the reference's
random_scalarizations.py (run_acquisition_function :75, ucb :160, thompson_sampling :264, EI :356,
random_scalarizations :509).  The arithmetic runs in the fused CUDA kernels (hm_rf_eval / hm_gp_mean_var +
hm_acq_score); the host only prepares the O(M) constants.
"""
import copy

import numpy as np

from . import _lib
from . import models
from .utility_functions import compute_data_array_scalarization, reciprocate_weights


def acquisition_constants(acquisition_function, scalarization_method, objectives, objective_weights, objective_limits,
                          data_array, iteration_number):
    """The per-call scalars the kernels need, computed exactly as the reference does:
      weights  : objective_weights, or reciprocate_weights(...) for modified_tchebyshev with EI/UCB/TS
                 (random_scalarizations.py:244, 336, 477)
      f_min    : EI only, 1 - (min(data[obj]) - lo)/diff with the limits widened by the observed data
                 (random_scalarizations.py:411-431)
      beta     : UCB only, sqrt(0.125 * log(2 * iteration + 1)) (random_scalarizations.py:187)
    """
    if scalarization_method not in _lib.SCALARIZATIONS:
        print("Error: unrecognized scalarization method:", scalarization_method)
        raise SystemExit
    w = reciprocate_weights(objective_weights) if scalarization_method == "modified_tchebyshev" else objective_weights
    weights = [w[o] for o in objectives]
    f_min = None
    beta = 0.0
    if acquisition_function == "EI":
        _, lim = compute_data_array_scalarization(data_array, objective_weights, copy.deepcopy(objective_limits),
                                                  scalarization_method)
        f_min = []
        for o in objectives:
            diff = lim[o][1] - lim[o][0]
            if diff == 0:
                diff = 1
            f_min.append(1 - (min(data_array[o]) - lim[o][0]) / diff)
    elif acquisition_function == "UCB":
        beta = float(np.sqrt(0.125 * np.log(2 * iteration_number + 1)))
    elif acquisition_function != "TS":
        print("Unrecognized acquisition function:", acquisition_function)
        raise SystemExit
    return weights, f_min, beta


def _feasibility_tensor(bufferx_enc32, classification_model, param_space):
    """P(feasible) per candidate from the feasibility classifier (random_scalarizations.py:395-409), or None."""
    if classification_model is None:
        return None
    return models.feasibility_probability_device(bufferx_enc32, classification_model, param_space)


def acquisition_scores_device(acquisition_function, candidates, objective_weights, regression_models, param_space,
                              scalarization_method, objective_limits, iteration_number, data_array, model_type,
                              classification_model=None):
    """Score candidates and leave the scores on the GPU: returns a CUDA fp64 tensor [N].
    candidates: list of dicts, list of tuples, a 2-D array of raw values [N][D], or an `EncodedCandidates`."""
    objectives = list(regression_models.keys())
    weights, f_min, beta = acquisition_constants(acquisition_function, scalarization_method, objectives,
                                                 objective_weights, objective_limits, data_array, iteration_number)
    params = _lib.make_acq_params(acquisition_function, scalarization_method, weights, f_min, beta)
    if model_type == "random_forest":
        from .forest import attach_packing, device_forest, rf_eval, rf_eval_packed
        forests = [device_forest(regression_models[o]) for o in objectives]
        if isinstance(candidates, models.EncodedCandidates) and candidates.codes is not None:
            # packed pool of an all-discrete space: the forest kernel walks the 64-bit codes (bit-identical results)
            feas = None
            if classification_model is not None:
                feas = _feasibility_tensor(candidates.get("float32"), classification_model, param_space)
            attach_packing(forests, candidates.device_space)
            return rf_eval_packed(forests, candidates.codes, acq=params, feas=feas)["score"]
        X32 = models.encode_to_device(candidates, param_space, "float32")
        feas = _feasibility_tensor(X32, classification_model, param_space)
        return rf_eval(forests, X32, acq=params, feas=feas)["score"]
    if model_type == "gaussian_process":
        import torch
        from .gp import device_gp, gp_mean_var
        X64 = models.encode_to_device(candidates, param_space, "float64")
        N = X64.shape[1]
        mean = torch.empty((len(objectives), N), dtype=torch.float64, device=X64.device)
        var = torch.empty_like(mean)
        for m, o in enumerate(objectives):
            gp_mean_var(device_gp(regression_models[o]), X64, mean[m], var[m])
        feas = None
        if classification_model is not None:
            feas = _feasibility_tensor(models.encode_to_device(candidates, param_space, "float32"), classification_model,
                                       param_space)
        return models.acq_score_device(params, mean, var, feas)
    print("Unrecognized model type:", model_type)
    raise SystemExit


def run_acquisition_function(acquisition_function, configurations, objective_weights, regression_models, param_space,
                             scalarization_method, objective_limits, iteration_number, data_array, model_type,
                             classification_model=None, number_of_cpus=0):
    """Drop-in for random_scalarizations.run_acquisition_function (:75-157): returns (list of values, [1]*N).
    This is the `optimization_function` plug-in of local_search (local_search.py:468-473)."""
    if acquisition_function not in _lib.ACQ_KINDS:
        print("Unrecognized acquisition function:", acquisition_function)
        raise SystemExit
    if model_type == "random_forest" and classification_model is None and not isinstance(configurations, models.EncodedCandidates):
        scalarized_values = _rf_scores_host(acquisition_function, configurations, objective_weights, regression_models,
                                            param_space, scalarization_method, objective_limits, iteration_number,
                                            data_array).tolist()
    else:
        scores = acquisition_scores_device(acquisition_function, configurations, objective_weights, regression_models,
                                           param_space, scalarization_method, objective_limits, iteration_number,
                                           data_array, model_type, classification_model)
        scalarized_values = scores.cpu().tolist()
    return scalarized_values, [1] * len(scalarized_values)


_host_scorers = {}


def _rf_scores_host(acquisition_function, configurations, objective_weights, regression_models, param_space,
                    scalarization_method, objective_limits, iteration_number, data_array):
    """Host in, host out through ONE C-ABI call (hm_rf_score_raw_host: H2D of the raw values -> hm_encode -> fused forest
    + acquisition kernel -> D2H, chunked and double-buffered): the latency path of the hill climb's small neighbour
    batches and the bandwidth path of large host-resident pools alike.  Returns a float64 numpy array [N]."""
    import numpy as np
    from .encoding import raw_matrix
    from .forest import device_forest
    from .scoring import HostScorer
    from .space_device import device_space_for
    _lib.require_cuda()
    objectives = list(regression_models.keys())
    weights, f_min, beta = acquisition_constants(acquisition_function, scalarization_method, objectives,
                                                 objective_weights, objective_limits, data_array, iteration_number)
    params = _lib.make_acq_params(acquisition_function, scalarization_method, weights, f_min, beta)
    raw = np.ascontiguousarray(raw_matrix(configurations, param_space))
    out = np.empty((raw.shape[0],), dtype=np.float64)
    if raw.shape[0] == 0:
        return out
    ds = device_space_for(param_space)
    scorer = _host_scorers.get(ds.n_encoded)
    if scorer is None:
        scorer = _host_scorers[ds.n_encoded] = HostScorer(ds.n_encoded, chunk=1 << 20)
    forests = [device_forest(regression_models[o]) for o in objectives]
    scorer.score_raw(ds, forests, raw, params, k=0, scores_out=out)
    return out


def _device_scores(candidates, **p):
    return acquisition_scores_device(p["acquisition_function"], candidates, p["objective_weights"],
                                     p["regression_models"], p["param_space"], p["scalarization_method"],
                                     p["objective_limits"], p["iteration_number"], p["data_array"], p["model_type"],
                                     p.get("classification_model"))


_device_scores.batch_invariant = True        # a candidate's score does not depend on the rest of the batch
run_acquisition_function.device_scores = _device_scores


def _exact_scores(rows, **p):
    """The reference's own value of run_acquisition_function for a HANDFUL of candidates (rows: [n][D] raw values, or
    an EncodedCandidates), used to order near-tied finalists exactly as the reference would (exact.py): the GPU
    reports their leaf ids (integers, bit-exact) and -- with a feasibility classifier -- P(feasible) (bit-exact with
    predict_proba); the Hutter tables, the variance with its libm pow and scipy's norm.cdf / pdf are evaluated on the
    host.  Random forests only (returns None otherwise: GPy's arithmetic cannot be reproduced bit for bit)."""
    if p["model_type"] != "random_forest":
        return None
    from . import exact
    from .forest import device_forest, rf_eval
    models_ = p["regression_models"]
    objectives = list(models_.keys())
    X32 = models.encode_to_device(rows, p["param_space"], "float32")
    out = rf_eval([device_forest(models_[o]) for o in objectives], X32, want_leaves=True)
    leaves = [lv.cpu().numpy() for lv in out["leaves"]]
    feas = None
    if p.get("classification_model") is not None:
        feas = _feasibility_tensor(X32, p["classification_model"], p["param_space"]).cpu().numpy()
    return exact.reference_rf_scores_from_leaves(p["acquisition_function"], p["scalarization_method"], models_, leaves,
                                                 p["objective_weights"], p["objective_limits"], p["iteration_number"],
                                                 p["data_array"], feas)


run_acquisition_function.exact_scores = _exact_scores


def _acq_from_bufferx(kind, bufferx, data_array, objective_weights, regression_models, param_space,
                      scalarization_method, objective_limits, iteration_number, model_type, classification_model):
    scores = acquisition_scores_device(kind, bufferx, objective_weights, regression_models, param_space,
                                       scalarization_method, objective_limits, iteration_number, data_array, model_type,
                                       classification_model)
    lim = copy.deepcopy(objective_limits)
    if kind == "EI":
        _, lim = compute_data_array_scalarization(data_array, objective_weights, lim, scalarization_method)
    return scores.cpu().numpy(), lim


def ucb(bufferx, objective_weights, regression_models, param_space, scalarization_method, objective_limits,
        iteration_number, model_type, classification_model=None, number_of_cpus=0):
    """random_scalarizations.py:160-261 -> (values ndarray, objective limits)"""
    return _acq_from_bufferx("UCB", bufferx, None, objective_weights, regression_models, param_space,
                             scalarization_method, objective_limits, iteration_number, model_type, classification_model)


def thompson_sampling(bufferx, objective_weights, regression_models, param_space, scalarization_method,
                      objective_limits, model_type, classification_model=None, number_of_cpus=0):
    """random_scalarizations.py:264-353.  RF: the 'sample' is the plain forest prediction (models.py:806-807)."""
    if model_type != "random_forest":
        raise _lib.HmError("TS with a GP is rewritten to EI by the reference (bo.py:211-216); not on the hot path")
    return _acq_from_bufferx("TS", bufferx, None, objective_weights, regression_models, param_space,
                             scalarization_method, objective_limits, 0, model_type, classification_model)


def EI(bufferx, data_array, objective_weights, regression_models, param_space, scalarization_method, objective_limits,
       iteration_number, model_type, classification_model=None, number_of_cpus=0):
    """random_scalarizations.py:356-506 -> (values ndarray, limits widened by the observed data)"""
    return _acq_from_bufferx("EI", bufferx, data_array, objective_weights, regression_models, param_space,
                             scalarization_method, objective_limits, iteration_number, model_type, classification_model)


def random_scalarizations(config, data_array, param_space, fast_addressing_of_data_array, regression_models,
                          iteration_number, objective_weights, objective_limits, classification_model=None,
                          profiling=None, acquisition_function_optimizer="local_search"):
    """random_scalarizations.py:509-598: pack the parameters and run the acquisition optimiser."""
    from .local_search import local_search
    p = {
        "regression_models": regression_models,
        "iteration_number": iteration_number,
        "data_array": data_array,
        "classification_model": classification_model,
        "param_space": param_space,
        "objective_weights": objective_weights,
        "model_type": config["models"]["model"],
        "objective_limits": objective_limits,
        "acquisition_function": config["acquisition_function"],
        "scalarization_method": config["scalarization_method"],
        "number_of_cpus": config["number_of_cpus"],
    }
    if acquisition_function_optimizer == "local_search":
        _, best_configuration = local_search(
            config["local_search_starting_points"], config["local_search_random_points"], param_space,
            fast_addressing_of_data_array, False, run_acquisition_function, p, config["scalarization_key"],
            config["number_of_cpus"], previous_points=data_array, profiling=profiling)
        return best_configuration
    if acquisition_function_optimizer == "cma_es":
        raise _lib.HmError("cma_es issues single-candidate, latency-bound calls on continuous spaces and is out of "
                           "scope of the batched hot path (SURVEY.md section 2); use the reference's cma_es with "
                           "hypermapper_b200.random_scalarizations.run_acquisition_function as its objective")
    print("Unrecognized acquisition function optimizer:", acquisition_function_optimizer)
    raise SystemExit
