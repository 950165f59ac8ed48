"""ctypes binding of the C-ABI in include/hm_b200.h (libhm_b200.so, built in-tree by csrc/Makefile).

There is no CPU fallback: if the library is missing or a call fails, an exception is raised.
torch is used for device memory and streams only.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# HM_B200_LIB selects another build of the same C-ABI (the checked build, experiment variants); default = the in-tree one
SO_PATH = os.environ.get("HM_B200_LIB") or os.path.join(CSRC, "libhm_b200.so")

HM_MAX_OBJECTIVES = 8
ACQ_KINDS = {"EI": 0, "UCB": 1, "TS": 2}
SCALARIZATIONS = {"linear": 0, "tchebyshev": 1, "modified_tchebyshev": 2}

# every symbol include/hm_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "hm_last_error", "hm_device_sm_count", "hm_version", "hm_fp64_peak_tflops", "hm_checked_failures",
    "hm_forest_create", "hm_forest_destroy", "hm_forest_n_trees", "hm_forest_n_features", "hm_forest_staged",
    "hm_forest_set_coherence", "hm_forest_n_groups", "hm_forest_leaf_resident",
    "hm_forest_attach_packing", "hm_forest_packed", "hm_forest_packed_groups", "hm_rf_eval_packed",
    "hm_rf_eval", "hm_acq_score",
    "hm_rf_score_topk_workspace_bytes", "hm_rf_score_topk", "hm_rf_score_topk_packed",
    "hm_topk_workspace_bytes", "hm_topk", "hm_topk_merge",
    "hm_gp_create", "hm_gp_destroy", "hm_gp_mean_var",
    "hm_pareto_workspace_bytes", "hm_pareto_filter", "hm_pareto_pass1", "hm_pareto_pass1_compact",
    "hm_space_create", "hm_space_destroy", "hm_space_n_params", "hm_space_n_encoded",
    "hm_space_packed_bits", "hm_space_packed_field", "hm_encode_packed", "hm_sample_pool_packed",
    "hm_rf_score_packed_host",
    "hm_encode", "hm_sample_pool", "hm_sample_rows", "hm_prior_prob", "hm_prior_weight",
    "hm_minmax", "hm_bopro_ratio", "hm_bopro_finish", "hm_ls_neighbors", "hm_ls_move",
    "hm_host_ctx_create", "hm_host_ctx_destroy", "hm_rf_score_host", "hm_rf_score_raw_host", "hm_gp_score_host",
]


class HmError(RuntimeError):
    pass


class AcqParams(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("scalarization", ctypes.c_int32),
        ("n_objectives", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("weight", ctypes.c_double * HM_MAX_OBJECTIVES),
        ("f_min", ctypes.c_double * HM_MAX_OBJECTIVES),
        ("beta", ctypes.c_double),
    ]


class BoproParams(ctypes.Structure):
    """hm_bopro_params_t"""
    _fields_ = [
        ("n_objectives", ctypes.c_int32),
        ("prior_mode", ctypes.c_int32),
        ("discrete_space", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("weight", ctypes.c_double * HM_MAX_OBJECTIVES),
        ("threshold", ctypes.c_double * HM_MAX_OBJECTIVES),
        ("prior_lo", ctypes.c_double),
        ("prior_hi", ctypes.c_double),
        ("posterior_floor", ctypes.c_double),
        ("discrete_divisor", ctypes.c_double),
        ("model_exponent", ctypes.c_double),
    ]


def build(verbose=False):
    """Compile every CUDA source for sm_100a into csrc/libhm_b200.so (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC, "-j", "8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise HmError("building libhm_b200.so failed")
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise HmError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)" % SO_PATH)
    L = ctypes.CDLL(SO_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    L.hm_last_error.restype = ctypes.c_char_p
    L.hm_device_sm_count.restype = i32
    L.hm_version.restype = i32
    L.hm_fp64_peak_tflops.restype = dbl
    L.hm_checked_failures.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.hm_forest_create.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(vp)]
    L.hm_forest_destroy.argtypes = [vp]
    L.hm_forest_destroy.restype = None
    for f in ("hm_forest_n_trees", "hm_forest_n_features", "hm_forest_staged", "hm_forest_n_groups"):
        getattr(L, f).argtypes = [vp]
    for f in ("hm_forest_packed", "hm_forest_packed_groups"):
        getattr(L, f).argtypes = [vp]
    L.hm_forest_leaf_resident.argtypes = [vp, i32]
    L.hm_forest_set_coherence.argtypes = [vp, i32]
    L.hm_forest_attach_packing.argtypes = [vp, vp]
    L.hm_rf_eval_packed.argtypes = [ctypes.POINTER(vp), i32, vp, i64, ctypes.POINTER(vp), vp, vp, vp,
                                    ctypes.POINTER(AcqParams), vp, vp, vp]
    L.hm_rf_eval.argtypes = [ctypes.POINTER(vp), i32, vp, i64, i64, ctypes.POINTER(vp), vp, vp, vp,
                             ctypes.POINTER(AcqParams), vp, vp, vp]
    L.hm_acq_score.argtypes = [ctypes.POINTER(AcqParams), vp, vp, i64, vp, i64, vp, vp]
    L.hm_rf_score_topk_workspace_bytes.argtypes = [i32]
    L.hm_rf_score_topk_workspace_bytes.restype = i64
    L.hm_rf_score_topk.argtypes = [ctypes.POINTER(vp), i32, vp, i64, i64, ctypes.POINTER(AcqParams), vp, i32, vp, vp, vp, vp]
    L.hm_rf_score_topk_packed.argtypes = [ctypes.POINTER(vp), i32, vp, i64, ctypes.POINTER(AcqParams), vp, i32, vp, vp,
                                          vp, vp]
    L.hm_topk_workspace_bytes.argtypes = [i32]
    L.hm_topk_workspace_bytes.restype = i64
    L.hm_topk.argtypes = [vp, i64, i64, i32, vp, vp, vp, vp]
    L.hm_topk_merge.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp]
    L.hm_gp_create.argtypes = [i32, i32, vp, vp, vp, vp, dbl, dbl, dbl, dbl, i32, ctypes.POINTER(vp)]
    L.hm_gp_destroy.argtypes = [vp]
    L.hm_gp_destroy.restype = None
    L.hm_gp_mean_var.argtypes = [vp, vp, i64, i64, vp, vp, vp]
    L.hm_pareto_workspace_bytes.argtypes = [i64, i32]
    L.hm_pareto_workspace_bytes.restype = i64
    L.hm_pareto_filter.argtypes = [vp, i64, i32, vp, vp, ctypes.POINTER(i64), vp]
    L.hm_pareto_pass1.argtypes = [vp, i64, i32, vp, vp, ctypes.POINTER(i64), vp]
    L.hm_pareto_pass1_compact.argtypes = [vp, i64, i32, vp, vp, vp, vp, i64, i64, ctypes.POINTER(i64),
                                          ctypes.POINTER(ctypes.c_int32), vp]
    u64 = ctypes.c_uint64
    L.hm_space_create.argtypes = [i32, vp, vp, i64, ctypes.POINTER(vp)]
    L.hm_space_destroy.argtypes = [vp]
    L.hm_space_destroy.restype = None
    L.hm_space_n_params.argtypes = [vp]
    L.hm_space_n_encoded.argtypes = [vp]
    L.hm_space_packed_bits.argtypes = [vp]
    L.hm_space_packed_field.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.hm_encode_packed.argtypes = [vp, vp, i32, i64, i64, vp, vp, vp]
    L.hm_sample_pool_packed.argtypes = [vp, u64, i64, i64, i32, vp, vp, i64, vp]
    L.hm_encode.argtypes = [vp, vp, i32, i64, i64, vp, i64, vp, i64, vp, vp]
    L.hm_sample_pool.argtypes = [vp, u64, i64, i64, i32, vp, i64, vp, i64, vp, i64, vp, i32, vp, vp]
    L.hm_sample_rows.argtypes = [vp, u64, i32, vp, i64, vp, vp]
    L.hm_prior_prob.argtypes = [vp, vp, i32, i64, i64, vp, i32, vp, vp, vp]
    L.hm_prior_weight.argtypes = [vp, vp, i64, dbl, dbl, i32, vp, vp, vp]
    L.hm_minmax.argtypes = [vp, i64, i32, vp, vp, vp]
    L.hm_bopro_ratio.argtypes = [ctypes.POINTER(BoproParams), vp, vp, vp, i64, i64, vp, vp]
    L.hm_bopro_finish.argtypes = [vp, i64, i32, dbl, dbl, dbl, vp, vp, vp, vp]
    L.hm_ls_neighbors.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp]
    L.hm_ls_move.argtypes = [vp, vp, i32, i32, vp, i32, vp, vp, vp, vp]
    L.hm_host_ctx_create.argtypes = [i64, i32, ctypes.POINTER(vp)]
    L.hm_host_ctx_destroy.argtypes = [vp]
    L.hm_host_ctx_destroy.restype = None
    L.hm_rf_score_host.argtypes = [vp, ctypes.POINTER(vp), i32, vp, i64, i64, ctypes.POINTER(AcqParams), vp, i32, vp, vp]
    L.hm_rf_score_packed_host.argtypes = [vp, ctypes.POINTER(vp), i32, vp, i64, ctypes.POINTER(AcqParams), vp, i32, vp, vp]
    L.hm_rf_score_raw_host.argtypes = [vp, vp, ctypes.POINTER(vp), i32, vp, i64, ctypes.POINTER(AcqParams), vp, i32, vp,
                                       vp, ctypes.POINTER(ctypes.c_int32)]
    L.hm_gp_score_host.argtypes = [vp, ctypes.POINTER(vp), i32, i32, vp, i64, i64, ctypes.POINTER(AcqParams), vp, i32, vp,
                                   vp]
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise HmError("%s failed (code %d): %s" % (what, rc, lib().hm_last_error().decode()))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise HmError("no CUDA device: hypermapper_b200 has no CPU fallback")
    return torch


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t):
    """device pointer of a torch tensor (or None)"""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def hptr(a):
    """host pointer of a contiguous numpy array (or None)"""
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


def make_acq_params(kind, scalarization, weights, f_min=None, beta=0.0):
    p = AcqParams()
    if kind not in ACQ_KINDS:
        print("Unrecognized acquisition function:", kind)
        raise SystemExit
    if scalarization not in SCALARIZATIONS:
        print("Error: unrecognized scalarization method:", scalarization)
        raise SystemExit
    p.kind = ACQ_KINDS[kind]
    p.scalarization = SCALARIZATIONS[scalarization]
    M = len(weights)
    if M > HM_MAX_OBJECTIVES:
        raise HmError("at most %d objectives are supported" % HM_MAX_OBJECTIVES)
    p.n_objectives = M
    for i in range(M):
        p.weight[i] = float(weights[i])
        p.f_min[i] = float(f_min[i]) if f_min is not None else 0.0
    p.beta = float(beta)
    return p
