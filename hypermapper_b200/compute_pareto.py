"""Pareto filter of the hot path, same entry points as the reference's compute_pareto.py
(sequential_is_pareto_efficient :89-117, _dumb :64-86, parallel_is_pareto_efficient :180-200,
compute_pareto :203-284).  The dominance sweep runs in the K4 kernels (hm_pareto_filter)."""
import csv
import ctypes
import multiprocessing as mp

import numpy as np

from . import _lib


def _as_cost_matrix(costs):
    costs = np.asarray(costs)
    if costs.dtype.kind in ("U", "S", "O"):
        costs = costs.astype(np.float64)        # compute_pareto.py:96-98 (string input)
    if costs.ndim != 2:
        raise ValueError("costs must be an (n_points, n_costs) array")
    return np.ascontiguousarray(costs, dtype=np.float64)


# workspaces of the last few (device, N, M) shapes: a BO loop filters the same shape over and over, and the library replays
# one CUDA graph per (costs, mask, workspace) triple -- a fresh workspace per call would also mean a fresh graph capture.
# Calls are stream-ordered on the current stream, so sequential calls may share a workspace.
_WS_CACHE = {}
_WS_CACHE_MAX = 2


def _workspace(torch, L, N, M, device):
    key = (device.index, int(N), int(M))
    ws = _WS_CACHE.get(key)
    if ws is None:
        while len(_WS_CACHE) >= _WS_CACHE_MAX:
            _WS_CACHE.pop(next(iter(_WS_CACHE)))
        ws = torch.empty((int(L.hm_pareto_workspace_bytes(N, M)),), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = ws
    return ws


def pareto_mask_device(costs_dev, pass2=True):
    """costs_dev: CUDA fp64 tensor [N][M] row-major -> CUDA uint8 mask [N]."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    assert costs_dev.is_cuda and costs_dev.dtype == torch.float64 and costs_dev.dim() == 2 and costs_dev.is_contiguous()
    N, M = costs_dev.shape
    mask = torch.empty((N,), dtype=torch.uint8, device=costs_dev.device)
    if N == 0:
        return mask
    if M > _lib.HM_MAX_OBJECTIVES:
        raise _lib.HmError("at most %d cost columns are supported" % _lib.HM_MAX_OBJECTIVES)
    ws = _workspace(torch, L, N, M, costs_dev.device)
    n_front = ctypes.c_int64(0)
    fn = L.hm_pareto_filter if pass2 else L.hm_pareto_pass1
    rc = fn(_lib.dptr(costs_dev), N, M, _lib.dptr(mask), _lib.dptr(ws), ctypes.byref(n_front), _lib.stream_ptr())
    _lib.check(rc, "hm_pareto_filter" if pass2 else "hm_pareto_pass1")
    return mask


def pareto_pass1_compact(costs_dev, cap, index_base=0, rows_out=None):
    """hm_pareto_pass1_compact: pass 1 on the device block + its survivors compacted in index order.
    -> (mask uint8 [N], rows fp64 [cap][M], idx int64 [cap], n_alive, saw_nan); rows / idx valid for j < min(n_alive, cap).
    rows_out (optional): a contiguous CUDA fp64 [cap][M] view the rows are written to (e.g. inside an exchange buffer)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    assert costs_dev.is_cuda and costs_dev.dtype == torch.float64 and costs_dev.dim() == 2 and costs_dev.is_contiguous()
    N, M = costs_dev.shape
    dev = costs_dev.device
    mask = torch.empty((N,), dtype=torch.uint8, device=dev)
    if rows_out is None:
        rows = torch.empty((cap, M), dtype=torch.float64, device=dev)
    else:
        rows = rows_out
        assert rows.is_cuda and rows.dtype == torch.float64 and rows.shape == (cap, M) and rows.is_contiguous()
    idx = torch.empty((cap,), dtype=torch.int64, device=dev)
    if N == 0:
        return mask, rows, idx, 0, False
    if M > _lib.HM_MAX_OBJECTIVES:
        raise _lib.HmError("at most %d cost columns are supported" % _lib.HM_MAX_OBJECTIVES)
    ws = _workspace(torch, L, N, M, dev)
    n_alive, saw_nan = ctypes.c_int64(0), ctypes.c_int32(0)
    rc = L.hm_pareto_pass1_compact(_lib.dptr(costs_dev), N, M, _lib.dptr(mask), _lib.dptr(ws), _lib.dptr(rows),
                                   _lib.dptr(idx), int(cap), int(index_base), ctypes.byref(n_alive),
                                   ctypes.byref(saw_nan), _lib.stream_ptr())
    _lib.check(rc, "hm_pareto_pass1_compact")
    return mask, rows, idx, int(n_alive.value), bool(saw_nan.value)


def sequential_is_pareto_efficient(costs):
    """:param costs: An (n_points, n_costs) array  :return: A (n_points, ) boolean array (compute_pareto.py:89-117)"""
    torch = _lib.require_cuda()
    c = _as_cost_matrix(costs)
    if c.shape[0] == 0:
        return np.ones(0, dtype=bool)
    mask = pareto_mask_device(torch.from_numpy(c).cuda(), pass2=True)
    return mask.cpu().numpy().astype(bool)


def sequential_is_pareto_efficient_dumb(costs):
    """compute_pareto.py:64-86.  Its first pass tests every point against ALL points instead of sweeping; for NaN-free
    input that is the same set as the sweep (strict dominance is transitive), which is what is computed here."""
    c = _as_cost_matrix(costs)
    if np.isnan(c).any():
        raise _lib.HmError("sequential_is_pareto_efficient_dumb with NaN costs is not supported on the GPU path")
    return sequential_is_pareto_efficient(c)


def _chunk_bounds(n, number_of_cpus):
    """domain decomposition of utility_functions.domain_decomposition_and_parallel_computation (:485-515)"""
    if number_of_cpus == 0:
        number_of_cpus = mp.cpu_count()
    if number_of_cpus > 8:
        number_of_cpus = 8
    chunks = min(number_of_cpus, n)
    f = float(chunks)
    bounds = [(int(n * (c / f)), int(n * ((c + 1) / f))) for c in range(chunks - 1)]
    bounds.append((int(n * ((chunks - 1) / f)), n))
    return bounds


def parallel_is_pareto_efficient(debug, predictions, costs, number_of_cpus=0):
    """compute_pareto.py:180-200: filter chunk-wise, keep the survivors, filter the survivors again.  `predictions` is
    reduced in place to the chunk survivors (as the reference does); returns the mask over those survivors."""
    c = _as_cost_matrix(costs)
    keep = np.concatenate([sequential_is_pareto_efficient(c[lo:hi]) for lo, hi in _chunk_bounds(len(c), number_of_cpus)])
    for k in predictions:
        predictions[k] = predictions[k][keep]
    return sequential_is_pareto_efficient(c[keep])


def compute_pareto(param_space, input_data_file, output_pareto_file, debug=False, number_of_cpus=0):
    """compute_pareto.py:203-284: CSV in -> drop infeasible rows -> Pareto of the FIRST TWO objectives -> CSV out.
    CSV parsing / typing is the reference Space's job (load_data_file, convert_types_to_string)."""
    objectives = param_space.get_optimization_parameters()
    x_select, y_select = objectives[0], objectives[1]
    feasible = param_space.get_feasible_parameter()[0]
    data_array, _ = param_space.load_data_file(input_data_file, debug, number_of_cpus=number_of_cpus)
    if feasible is not None:
        keep_rows = [i for i, v in enumerate(data_array[feasible]) if not (v == False)]  # noqa: E712  (sic, :242)
        for key in list(data_array.keys()):
            col = data_array[key]
            data_array[key] = [col[i] for i in keep_rows]
        if len(data_array[feasible]) == 0:
            print("Warning: after removing the non-valid rows in file %s the data array is now empty." % input_data_file)
            with open(output_pareto_file, "w") as f:
                csv.writer(f).writerow(list(data_array.keys()))
            return 0
    costs = np.column_stack((data_array[x_select], data_array[y_select]))
    mask = sequential_is_pareto_efficient(costs)
    count = 0
    with open(output_pareto_file, "w") as f:
        w = csv.writer(f)
        w.writerow(list(data_array.keys()))
        cols = [param_space.convert_types_to_string(j, data_array) for j in list(data_array.keys())]
        rows = list(zip(*cols))
        for i in range(len(mask)):
            if mask[i]:
                w.writerow(rows[i])
                count += 1
    return count
