"""Host-buffer scoring: the call a reference-side binding makes when candidates live in host memory.
Wraps hm_host_ctx_* / hm_rf_score_host (include/hm_b200.h): chunked, double-buffered H2D -> fused forest +
acquisition kernel -> D2H, plus the stable device top-k."""
import ctypes

import numpy as np

from . import _lib


class HostScorer:
    def __init__(self, max_features, chunk=1 << 21):
        _lib.require_cuda()
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().hm_host_ctx_create(int(chunk), int(max_features), ctypes.byref(h)), "hm_host_ctx_create")
        self.handle = h

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().hm_host_ctx_destroy(h)
            except Exception:
                pass
            self.handle = None

    def score(self, forests, x_host, acq, k=0, scores_out=None):
        """forests: list of DeviceForest; x_host: float32 column-major [D_enc][N] host array (numpy, or a pinned torch
        CPU tensor for full-speed async copies); acq: AcqParams.  scores_out: optional float64 [N] host buffer.
        Returns (scores_out, topk_values[k], topk_indices[k])."""
        if hasattr(x_host, "data_ptr"):
            D, N = x_host.shape
            assert x_host.is_contiguous() and not x_host.is_cuda and str(x_host.dtype) == "torch.float32"
            xp = ctypes.c_void_p(x_host.data_ptr())
        else:
            assert x_host.dtype == np.float32 and x_host.flags.c_contiguous
            D, N = x_host.shape
            xp = _lib.hptr(x_host)
        sp = None
        if scores_out is not None:
            sp = ctypes.c_void_p(scores_out.data_ptr()) if hasattr(scores_out, "data_ptr") else _lib.hptr(scores_out)
        tv = np.empty(max(k, 1), dtype=np.float64)
        ti = np.empty(max(k, 1), dtype=np.int64)
        M = len(forests)
        handles = (ctypes.c_void_p * M)(*[f.handle for f in forests])
        rc = _lib.lib().hm_rf_score_host(self.handle, handles, M, xp, N, N, ctypes.byref(acq), sp, int(k),
                                         _lib.hptr(tv), _lib.hptr(ti))
        _lib.check(rc, "hm_rf_score_host")
        return scores_out, tv[:k], ti[:k]

    def score_packed(self, forests, codes_host, acq, k=0, scores_out=None):
        """Like score(), from packed 64-bit codes of an all-discrete space (DeviceSpace.pack_codes_host): codes_host is a
        uint64 / int64 [N] host array (numpy, or a pinned torch CPU tensor); 8 bytes per candidate cross PCIe.  The
        forests must have been through forest.attach_packing."""
        if hasattr(codes_host, "data_ptr"):
            N = codes_host.shape[0]
            assert codes_host.is_contiguous() and not codes_host.is_cuda and codes_host.element_size() == 8
            cp = ctypes.c_void_p(codes_host.data_ptr())
        else:
            assert codes_host.dtype in (np.uint64, np.int64) and codes_host.flags.c_contiguous
            N = codes_host.shape[0]
            cp = _lib.hptr(codes_host)
        sp = None
        if scores_out is not None:
            sp = ctypes.c_void_p(scores_out.data_ptr()) if hasattr(scores_out, "data_ptr") else _lib.hptr(scores_out)
        tv = np.empty(max(k, 1), dtype=np.float64)
        ti = np.empty(max(k, 1), dtype=np.int64)
        M = len(forests)
        handles = (ctypes.c_void_p * M)(*[f.handle for f in forests])
        rc = _lib.lib().hm_rf_score_packed_host(self.handle, handles, M, cp, N, ctypes.byref(acq), sp, int(k),
                                                _lib.hptr(tv), _lib.hptr(ti))
        _lib.check(rc, "hm_rf_score_packed_host")
        return scores_out, tv[:k], ti[:k]

    def score_raw(self, device_space, forests, raw_host, acq, k=0, scores_out=None):
        """Like score(), from RAW parameter values: raw_host is a float64 row-major [N][D] host array (numpy, or a pinned
        torch CPU tensor) in input-parameter order, categoricals as indices.  Raises ValueError like the reference when
        a discrete value is not one of its parameter's values."""
        if hasattr(raw_host, "data_ptr"):
            N, D = raw_host.shape
            assert raw_host.is_contiguous() and not raw_host.is_cuda and str(raw_host.dtype) == "torch.float64"
            rp = ctypes.c_void_p(raw_host.data_ptr())
        else:
            assert raw_host.dtype == np.float64 and raw_host.flags.c_contiguous
            N, D = raw_host.shape
            rp = _lib.hptr(raw_host)
        assert D == device_space.D
        sp = None
        if scores_out is not None:
            sp = ctypes.c_void_p(scores_out.data_ptr()) if hasattr(scores_out, "data_ptr") else _lib.hptr(scores_out)
        tv = np.empty(max(k, 1), dtype=np.float64)
        ti = np.empty(max(k, 1), dtype=np.int64)
        M = len(forests)
        handles = (ctypes.c_void_p * M)(*[f.handle for f in forests])
        bad = ctypes.c_int32(0)
        rc = _lib.lib().hm_rf_score_raw_host(self.handle, device_space.handle, handles, M, rp, N, ctypes.byref(acq), sp,
                                             int(k), _lib.hptr(tv), _lib.hptr(ti), ctypes.byref(bad))
        _lib.check(rc, "hm_rf_score_raw_host")
        if bad.value:
            raise ValueError("a discrete parameter value is not in its list of values (models.py:962-983)")
        return scores_out, tv[:k], ti[:k]

    def score_gp(self, device_gps, x_host, acq, k=0, scores_out=None):
        """GP models: x_host is a float64 column-major [D_enc][N] host array (numpy or pinned torch CPU tensor) of encoded
        candidates; device_gps: list of DeviceGP, one per objective.  Returns (scores_out, topk_values, topk_indices)."""
        if hasattr(x_host, "data_ptr"):
            D, N = x_host.shape
            assert x_host.is_contiguous() and not x_host.is_cuda and str(x_host.dtype) == "torch.float64"
            xp = ctypes.c_void_p(x_host.data_ptr())
        else:
            assert x_host.dtype == np.float64 and x_host.flags.c_contiguous
            D, N = x_host.shape
            xp = _lib.hptr(x_host)
        sp = None
        if scores_out is not None:
            sp = ctypes.c_void_p(scores_out.data_ptr()) if hasattr(scores_out, "data_ptr") else _lib.hptr(scores_out)
        tv = np.empty(max(k, 1), dtype=np.float64)
        ti = np.empty(max(k, 1), dtype=np.int64)
        M = len(device_gps)
        handles = (ctypes.c_void_p * M)(*[g.handle for g in device_gps])
        rc = _lib.lib().hm_gp_score_host(self.handle, handles, M, D, xp, N, N, ctypes.byref(acq), sp, int(k),
                                         _lib.hptr(tv), _lib.hptr(ti))
        _lib.check(rc, "hm_gp_score_host")
        return scores_out, tv[:k], ti[:k]
