"""Prediction entry points of the hot path, same names / signatures / return types as the reference's models.py:

    compute_model_mean_and_uncertainty            models.py:742-764
    compute_rf_prediction_mean_and_uncertainty    models.py:767-794
    compute_gp_prediction_mean_and_uncertainty    models.py:997-1024
    model_prediction                              models.py:879-890
    model_probabilities                           models.py:893-904
    preprocess_data_buffer                        models.py:921-938
    RFModel.get_leaves_per_sample / compute_rf_prediction / compute_rf_prediction_variance   models.py:174-275
      (module-level functions taking the model first; hypermapper_b200.install() binds them as methods)

`bufferx` is a list of tuples in input-parameter order (categoricals as integer indices), as in the reference;
a 2-D numpy array [N][D] of raw values, a list of dicts, or an `EncodedCandidates` are accepted as well (the
array forms skip the per-candidate Python objects that make 10^7 candidates infeasible in the reference).
"""
import sys

import numpy as np

from . import _lib
from .encoding import encoder_for, raw_matrix


class EncodedCandidates:
    """Candidates that are already encoded and resident on the GPU (column-major [D_enc][N]); `raw_cm` optionally
    holds their raw parameter values (column-major [D][N] fp64) for the prior-weighted acquisitions."""

    def __init__(self, x32=None, x64=None, raw_cm=None, codes=None, device_space=None):
        self.x32 = x32
        self.x64 = x64
        self.raw_cm = raw_cm
        self.codes = codes                  # packed 64-bit codes (all-discrete spaces, csrc/hm_pack.cuh) as int64 [N]
        self.device_space = device_space    # the DeviceSpace the codes / raw values belong to

    def __len__(self):
        if self.codes is not None:
            return int(self.codes.shape[0])
        t = self.x32 if self.x32 is not None else self.x64
        return int(t.shape[1])

    def get(self, dtype):
        torch = _lib.require_cuda()
        if self.x32 is None and self.x64 is None:
            # a packed pool: the encoded matrix is only built when something other than the forest kernel asks for it
            # (a feasibility classifier, a GP)
            x32, x64 = self.device_space.encode(self.raw_cm, want32=(dtype == "float32"), want64=(dtype != "float32"),
                                                col_major=True)
            self.x32, self.x64 = x32, x64
        if dtype == "float32":
            if self.x32 is None:
                self.x32 = self.x64.to(torch.float32)
            return self.x32
        if self.x64 is None:
            self.x64 = self.x32.to(torch.float64)
        return self.x64


def encode_host(bufferx, param_space, dtype=np.float64):
    """[D_enc][N] numpy array in the reference's encoded column order."""
    enc = encoder_for(param_space)
    if len(bufferx) > 0 and isinstance(bufferx, (list, tuple)) and isinstance(bufferx[0], dict):
        return enc.encode_configurations(bufferx, dtype=dtype)
    return enc.encode(bufferx, dtype=dtype)


def encode_to_device(bufferx, param_space, dtype):
    """models.preprocess_data_array (models.py:941-984) on the device (K5 hm_encode): the raw [N][D] values are copied
    to the GPU and encoded there, column-major.  float64 -> float32 is the round-to-nearest cast sklearn's
    check_array(dtype=float32) applies."""
    _lib.require_cuda()
    if isinstance(bufferx, EncodedCandidates):
        return bufferx.get(dtype)
    from .space_device import device_space_for
    raw = bufferx if (hasattr(bufferx, "is_cuda") and bufferx.is_cuda) else raw_matrix(bufferx, param_space)
    x32, x64 = device_space_for(param_space).encode(raw, want32=(dtype == "float32"), want64=(dtype != "float32"))
    return x32 if dtype == "float32" else x64


def mean_and_variance_device(candidates, regression_models, model_type, param_space):
    """Per-objective posterior mean / variance left on the GPU: CUDA fp64 tensors [M][N] (objectives in
    regression_models order).  NaN means are NOT yet patched to sys.maxsize (models.py:762-763); consumers do it."""
    torch = _lib.require_cuda()
    objectives = list(regression_models.keys())
    if model_type == "random_forest":
        from .forest import device_forest, rf_eval
        X32 = encode_to_device(candidates, param_space, "float32")
        out = rf_eval([device_forest(regression_models[o]) for o in objectives], X32, want_mean=True, want_var=True)
        return out["mean"], out["var"]
    if model_type == "gaussian_process":
        from .gp import device_gp, gp_mean_var
        X64 = encode_to_device(candidates, param_space, "float64")
        N = X64.shape[1]
        mean = torch.empty((len(objectives), N), dtype=torch.float64, device=X64.device)
        var = torch.empty_like(mean)
        for m, o in enumerate(objectives):
            gp_mean_var(device_gp(regression_models[o]), X64, mean[m], var[m])
        return mean, var
    print("Unrecognized model type:", model_type)
    raise SystemExit


def preprocess_data_buffer(bufferx, param_space):
    """models.py:921-938: list of encoded tuples."""
    host = encode_host(bufferx, param_space)
    return [tuple(row) for row in host.T]


def acq_score_device(params, mean, var, feas=None):
    """hm_acq_score on CUDA tensors mean/var [M][N] -> scores [N]."""
    import ctypes
    torch = _lib.require_cuda()
    M, N = mean.shape
    out = torch.empty((N,), dtype=torch.float64, device=mean.device)
    if N == 0:
        return out
    rc = _lib.lib().hm_acq_score(ctypes.byref(params), _lib.dptr(mean), _lib.dptr(var), N, _lib.dptr(feas), N,
                                 _lib.dptr(out), _lib.stream_ptr())
    _lib.check(rc, "hm_acq_score")
    return out


# ---- random forest ---------------------------------------------------------------------------------------
def compute_rf_prediction_mean_and_uncertainty(bufferx, model, param_space, var=False):
    from .forest import device_forest, rf_eval
    objectives = list(model.keys())
    X32 = encode_to_device(bufferx, param_space, "float32")
    out = rf_eval([device_forest(model[o]) for o in objectives], X32, want_mean=True, want_var=True)
    mean = out["mean"].cpu().numpy()
    variance = out["var"].cpu().numpy()
    means, unc = {}, {}
    for m, o in enumerate(objectives):
        means[o] = mean[m].copy()
        unc[o] = variance[m].copy() if var else np.sqrt(variance[m])
    return means, unc


def rf_get_leaves_per_sample(model, bufferx, param_space=None):
    """RFModel.get_leaves_per_sample (models.py:174-182): bufferx is ALREADY encoded (list of tuples or [N][D_enc]
    array); returns int64 [T][N]."""
    from .forest import DeviceForest, _NoHutter, forest_tables, rf_eval
    torch = _lib.require_cuda()
    X = np.asarray(bufferx, dtype=np.float64)
    X32 = torch.from_numpy(np.ascontiguousarray(X.astype(np.float32).T)).cuda()
    # an UNCACHED, table-less forest: the reference's fit_RF calls this in the middle of the fit (models.py:200),
    # before the Hutter tables exist and before the split thresholds are re-drawn in place
    out = rf_eval([DeviceForest(forest_tables(_NoHutter(model)))], X32, want_leaves=True)
    return out["leaves"][0].cpu().numpy().astype(np.int64)


def _leaf_tables(model):
    T = len(model.estimators_)
    return T, model.tree_means_per_leaf, model.tree_vars_per_leaf


def rf_compute_rf_prediction(model, leaf_per_sample):
    """RFModel.compute_rf_prediction (models.py:225-240) from a leaf matrix.  Host gather: this entry point takes a
    host leaf matrix in the reference API; the GPU path never materialises it (hm_rf_eval reduces in-register)."""
    T, means, _ = _leaf_tables(model)
    leaf = np.asarray(leaf_per_sample)
    out = np.zeros(leaf.shape[1])
    for t in range(T):
        tab = means[t]
        out += np.array([tab[l] for l in leaf[t]]) / T
    return out


def rf_compute_rf_prediction_variance(model, leaf_per_sample, sample_means):
    """RFModel.compute_rf_prediction_variance (models.py:242-275) from a leaf matrix (host, see above)."""
    T, means, variances = _leaf_tables(model)
    leaf = np.asarray(leaf_per_sample)
    n = leaf.shape[1]
    mov = np.zeros(n)
    vom = np.zeros(n)
    for t in range(T):
        mov += np.array([variances[t][l] for l in leaf[t]]) / T
        vom += np.array([means[t][l] ** 2 for l in leaf[t]]) / T
    vom = np.abs(vom - np.array([m ** 2 for m in np.asarray(sample_means, dtype=np.float64)]))
    out = mov + vom
    out[out == 0] = 0.00001
    return out


def model_prediction(bufferx, model, param_space):
    """models.py:879-890: plain forest prediction per objective (what thompson sampling consumes)."""
    from .forest import device_forest, rf_eval
    objectives = list(model.keys())
    X32 = encode_to_device(bufferx, param_space, "float32")
    out = rf_eval([device_forest(model[o]) for o in objectives], X32, want_pred=True)
    pred = out["pred"].cpu().numpy()
    return {o: pred[m].copy() for m, o in enumerate(objectives)}


def feasibility_probability_device(X32, classification_model, param_space):
    """P(True) of the feasibility classifier (random_scalarizations.py:395-409) as a CUDA fp64 tensor."""
    from .forest import classifier_forest, rf_eval
    feasible_parameter = param_space.get_feasible_parameter()[0]
    clf = classification_model[feasible_parameter]
    df = classifier_forest(clf, True)
    return rf_eval([df], X32, want_pred=True)["pred"][0]


def model_probabilities(bufferx, model, param_space):
    """models.py:893-904: predict_proba per feasibility output -> {name: ndarray [N][n_classes]}"""
    from .forest import classifier_forest, rf_eval
    X32 = encode_to_device(bufferx, param_space, "float32")
    out = {}
    for name, clf in model.items():
        cols = [rf_eval([classifier_forest(clf, c)], X32, want_pred=True)["pred"][0].cpu().numpy() for c in clf.classes_]
        out[name] = np.stack(cols, axis=1)
    return out


# ---- gaussian process --------------------------------------------------------------------------------------
def compute_gp_prediction_mean_and_uncertainty(bufferx, model, param_space, var=False):
    import torch
    from .gp import device_gp, gp_mean_var
    X64 = encode_to_device(bufferx, param_space, "float64")
    N = X64.shape[1]
    means, unc = {}, {}
    for o in model:
        mean = torch.empty((N,), dtype=torch.float64, device=X64.device)
        variance = torch.empty_like(mean)
        gp_mean_var(device_gp(model[o]), X64, mean, variance)
        means[o] = mean.cpu().numpy()
        v = variance.cpu().numpy()
        unc[o] = v if var else np.sqrt(v)
    return means, unc


def compute_model_mean_and_uncertainty(bufferx, model, model_type, param_space, var=False):
    if model_type == "random_forest":
        mean, uncertainty = compute_rf_prediction_mean_and_uncertainty(bufferx, model, param_space, var=var)
    elif model_type == "gaussian_process":
        mean, uncertainty = compute_gp_prediction_mean_and_uncertainty(bufferx, model, param_space, var=var)
    else:
        raise UnboundLocalError("cannot access local variable 'mean' where it is not associated with a value")
    for objective in mean:
        mean[objective][np.isnan(mean[objective])] = sys.maxsize
    return mean, uncertainty
