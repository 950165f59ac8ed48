"""Prior-weighted acquisition (piBO), same names and signatures as the reference's pibo.py:

    prior_weighted_acquisition_function     pibo.py:68-113     acquisition kernels + K7 hm_prior_prob + hm_prior_weight
    run_pibo                                pibo.py:116-213
"""
from . import _lib
from .prior_optimization import prior_probability_device, _raw_matrix


def prior_weighted_scores_device(configurations, param_space, acquisition_function_wrapper, acquisition_function,
                                 objective_weights, objective_limits, iteration_number, prior_beta, regression_models,
                                 classification_model, data_array, model_type, scalarization_method,
                                 prior_floor=10 ** -6):
    """acq(x) * (p(x) + prior_floor) ** (prior_beta / iteration_number), left on the GPU (CUDA fp64 tensor [N]).
    The reference calls the wrapper with classification_model=None (pibo.py:88-99) and, for TS / UCB, collapses the
    acquisition values to their minimum before weighting (pibo.py:103-104); both are reproduced."""
    torch = _lib.require_cuda()
    from . import random_scalarizations as rs
    from .models import EncodedCandidates
    raw = configurations if isinstance(configurations, EncodedCandidates) else _raw_matrix(configurations, param_space)
    if (acquisition_function_wrapper is rs.run_acquisition_function or acquisition_function_wrapper is None or
            isinstance(raw, EncodedCandidates)):
        acq = rs.acquisition_scores_device(acquisition_function, raw, objective_weights, regression_models,
                                           param_space, scalarization_method, objective_limits, iteration_number,
                                           data_array, model_type, None)
    else:
        inputs = list(param_space.get_input_parameters())
        confs = configurations if (len(configurations) and isinstance(configurations[0], dict)) else \
            [{p: row[j] for j, p in enumerate(inputs)} for row in raw]
        values, _ = acquisition_function_wrapper(acquisition_function, confs, objective_weights, regression_models,
                                                 param_space, scalarization_method, objective_limits,
                                                 iteration_number, data_array, model_type, classification_model=None,
                                                 number_of_cpus=0)
        acq = torch.as_tensor(values, dtype=torch.float64, device="cuda")
    prob = prior_probability_device(raw, param_space, objective_weights)
    N = acq.numel()
    out = torch.empty((N,), dtype=torch.float64, device="cuda")
    ws = torch.empty((1,), dtype=torch.int64, device="cuda")
    collapse = 1 if acquisition_function in ("TS", "UCB") else 0
    rc = _lib.lib().hm_prior_weight(_lib.dptr(acq), _lib.dptr(prob), N, float(prior_floor),
                                    prior_beta / iteration_number, collapse, _lib.dptr(out), _lib.dptr(ws),
                                    _lib.stream_ptr())
    _lib.check(rc, "hm_prior_weight")
    return out


def prior_weighted_acquisition_function(configurations, param_space, acquisition_function_wrapper,
                                        acquisition_function, objective_weights, objective_limits, iteration_number,
                                        prior_beta, regression_models, classification_model, data_array, model_type,
                                        scalarization_method, prior_floor=10 ** -6):
    """pibo.py:68-113 -> (list of values, feasibility indicators [1]*N)."""
    out = prior_weighted_scores_device(configurations, param_space, acquisition_function_wrapper,
                                       acquisition_function, objective_weights, objective_limits, iteration_number,
                                       prior_beta, regression_models, classification_model, data_array, model_type,
                                       scalarization_method, prior_floor)
    values = out.cpu().tolist()
    return values, [1] * len(values)


def _device_scores(candidates, **p):
    p = {k: v for k, v in p.items() if k != "number_of_cpus"}
    return prior_weighted_scores_device(candidates, **p)


prior_weighted_acquisition_function.device_scores = _device_scores


def run_pibo(config, data_array, param_space, fast_addressing_of_data_array, regression_models, iteration_number,
             objective_weights, objective_limits, classification_model=None, profiling=None,
             acquisition_function_optimizer="local_search"):
    """pibo.py:116-213: pack the parameters and run the acquisition optimiser."""
    from .local_search import local_search
    from .random_scalarizations import run_acquisition_function
    fp = {
        "param_space": param_space,
        "iteration_number": iteration_number,
        "regression_models": regression_models,
        "classification_model": classification_model,
        "objective_weights": objective_weights,
        "objective_limits": objective_limits,
        "scalarization_method": config["scalarization_method"],
        "model_type": config["models"]["model"],
        "acquisition_function": config["acquisition_function"],
        "acquisition_function_wrapper": run_acquisition_function,
        "data_array": data_array,
        "prior_beta": config["optimization_iterations"] * 0.1 if config["prior_beta"] == -1 else config["prior_beta"],
        "prior_floor": config["prior_floor"],
    }
    if acquisition_function_optimizer == "local_search":
        _, best = local_search(config["local_search_starting_points"], config["local_search_random_points"],
                               param_space, fast_addressing_of_data_array, False, prior_weighted_acquisition_function,
                               fp, config["scalarization_key"], config["number_of_cpus"], previous_points=data_array,
                               profiling=profiling)
        return best
    if acquisition_function_optimizer == "cma_es":
        raise _lib.HmError("cma_es is out of scope of the batched hot path (SURVEY.md section 2); use the reference's "
                           "cma_es with hypermapper_b200.pibo.prior_weighted_acquisition_function as its objective")
    print("Unrecognized acquisition function optimizer:", acquisition_function_optimizer)
    raise SystemExit
