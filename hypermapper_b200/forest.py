"""Fitted forest -> device-resident structure-of-arrays forest (hm_forest_create) and the launch wrappers.

Model fitting is out of scope (stays with sklearn / the reference's RFModel.fit_RF, models.py:184-223); this
module only reads the fitted model: sklearn's tree arrays plus the Hutter per-leaf tables
(tree_means_per_leaf / tree_vars_per_leaf, models.py:79-122).
"""
import ctypes
import weakref

import numpy as np

from . import _lib


def forest_tables(model, value_column=0, normalize_value=False):
    """Flatten `model` (reference RFModel, or any sklearn forest) into the host arrays hm_forest_create takes.

    The three per-leaf fp64 tables are computed here with the reference's own scalar arithmetic so that the
    device only has to add them up in tree order:
        q = m / T                (models.py:237-239)
        s = (m ** 2) / T         (models.py:262-264; `**` on a float scalar is libm pow, not m*m)
        u = v / T                (models.py:259-261)
    """
    ests = model.estimators_
    T = len(ests)
    means = getattr(model, "tree_means_per_leaf", None) or None
    vars_ = getattr(model, "tree_vars_per_leaf", None) or None
    offs = np.zeros(T + 1, dtype=np.int64)
    for t, est in enumerate(ests):
        offs[t + 1] = offs[t] + est.tree_.node_count
    n = int(offs[-1])
    left = np.empty(n, dtype=np.int32)
    right = np.empty(n, dtype=np.int32)
    feature = np.empty(n, dtype=np.int32)
    threshold = np.empty(n, dtype=np.float64)
    value = np.empty(n, dtype=np.float64)
    q = s = u = None
    if means is not None:
        q = np.zeros(n, dtype=np.float64)
        s = np.zeros(n, dtype=np.float64)
        u = np.zeros(n, dtype=np.float64)
    for t, est in enumerate(ests):
        tr = est.tree_
        b, e = int(offs[t]), int(offs[t + 1])
        left[b:e] = tr.children_left
        right[b:e] = tr.children_right
        feature[b:e] = tr.feature
        threshold[b:e] = tr.threshold
        val = np.asarray(tr.value, dtype=np.float64).reshape(e - b, -1)
        if normalize_value:
            # class fractions whatever the scikit-learn version: tree_.value holds weighted class COUNTS before 1.4
            # and fractions from 1.4 on; predict_proba normalises per node (sklearn/tree/_classes.py), so do the same
            tot = val.sum(axis=1)
            tot[tot == 0.0] = 1.0
            value[b:e] = val[:, value_column] / tot
        else:
            value[b:e] = val[:, value_column]
        if means is not None:
            for leaf, m in means[t].items():
                q[b + int(leaf)] = m / T
                s[b + int(leaf)] = (m ** 2) / T
            for leaf, v in vars_[t].items():
                u[b + int(leaf)] = v / T
    return {"T": T, "n_features": int(model.n_features_in_), "tree_off": offs, "left": left, "right": right,
            "feature": feature, "threshold": threshold, "q": q, "s": s, "u": u, "value": value}


def tables_from_arrays(tree_off, left, right, feature, threshold, leaf_mean, leaf_var, value, n_features):
    """Same as forest_tables but from already-flattened arrays (golden fixtures / synthetic workloads):
    leaf_mean / leaf_var are per-node arrays holding the Hutter tables at leaf positions."""
    T = len(tree_off) - 1
    q = s = u = None
    if leaf_mean is not None:
        lm = np.asarray(leaf_mean, dtype=np.float64)
        is_leaf = np.asarray(left) < 0
        q = np.where(is_leaf, lm / T, 0.0)
        s = np.zeros_like(lm)
        for i in np.nonzero(is_leaf)[0]:
            s[i] = (float(lm[i]) ** 2) / T      # scalar pow, as the reference
        u = np.where(is_leaf, np.asarray(leaf_var, dtype=np.float64) / T, 0.0)
    return {"T": T, "n_features": int(n_features), "tree_off": np.asarray(tree_off, dtype=np.int64),
            "left": np.asarray(left, dtype=np.int32), "right": np.asarray(right, dtype=np.int32),
            "feature": np.asarray(feature, dtype=np.int32), "threshold": np.asarray(threshold, dtype=np.float64),
            "q": q, "s": s, "u": u, "value": None if value is None else np.asarray(value, dtype=np.float64)}


class DeviceForest:
    """Owns one hm_forest handle on the current CUDA device."""

    def __init__(self, tables):
        _lib.require_cuda()
        L = _lib.lib()
        c = np.ascontiguousarray
        self._keep = [c(tables["tree_off"], dtype=np.int64), c(tables["feature"], dtype=np.int32),
                      c(tables["threshold"], dtype=np.float64), c(tables["left"], dtype=np.int32),
                      c(tables["right"], dtype=np.int32)]
        opt = [None if tables.get(k) is None else c(tables[k], dtype=np.float64) for k in ("q", "s", "u", "value")]
        h = ctypes.c_void_p()
        rc = L.hm_forest_create(int(tables["T"]), int(tables["n_features"]), *[_lib.hptr(a) for a in self._keep],
                                *[_lib.hptr(a) for a in opt], ctypes.byref(h))
        _lib.check(rc, "hm_forest_create")
        self.handle = h
        self.n_trees = int(tables["T"])
        self.n_features = int(tables["n_features"])
        self.has_hutter_tables = tables.get("q") is not None
        self.staged = bool(L.hm_forest_staged(h))
        self.n_groups = int(L.hm_forest_n_groups(h))
        self._keep = None

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().hm_forest_destroy(h)
            except Exception:
                pass
            self.handle = None


_cache = weakref.WeakKeyDictionary()


def _model_stamp(model):
    """Identity of what a DeviceForest was flattened from.  A new fit creates new estimator objects; the reference's
    fit_RF (models.py:184-223) additionally fills the Hutter tables and rewrites tree_.threshold IN PLACE after the
    sklearn fit, so the stamp also covers the tables (identity + length) and the threshold bytes of the first, middle and
    last tree (fit_RF re-draws the splits of every tree; three crc32s keep the check off the hill climb's latency path)."""
    import zlib
    ests = model.estimators_
    means = getattr(model, "tree_means_per_leaf", None)
    crc = 0
    for t in sorted({0, len(ests) // 2, len(ests) - 1}):
        crc = zlib.crc32(np.ascontiguousarray(ests[t].tree_.threshold).view(np.uint8), crc)
    return (id(ests), len(ests), id(ests[0].tree_), id(ests[-1].tree_), id(means), len(means) if means else 0, crc)


def device_forest(model):
    """DeviceForest for a fitted model, cached on the model object and rebuilt whenever the model's trees, thresholds or
    Hutter tables changed (see _model_stamp)."""
    if isinstance(model, DeviceForest):
        return model
    stamp = _model_stamp(model)
    try:
        hit = _cache.get(model)
    except TypeError:
        hit = None
    if hit is not None and hit[0] == stamp:
        return hit[1]
    df = DeviceForest(forest_tables(model))
    try:
        _cache[model] = (stamp, df)
    except TypeError:
        pass
    return df


_clf_cache = weakref.WeakKeyDictionary()


def classifier_forest(clf, cls):
    """Feasibility classifier (RandomForestClassifier, models.py:574-578) as a traversal forest whose leaf payload is
    the class fraction of `cls`: predict_proba = sum_t tree_.value[leaf, 0, class] / T (sklearn/ensemble/_forest.py),
    which is exactly the `pred` output of hm_rf_eval."""
    col = list(clf.classes_).index(cls)
    ests = clf.estimators_
    stamp = (id(ests), len(ests), id(ests[0].tree_), col)
    hit = _clf_cache.get(clf)
    if hit is not None and stamp in hit:
        return hit[stamp]
    df = DeviceForest(forest_tables(_NoHutter(clf), value_column=col, normalize_value=True))
    _clf_cache.setdefault(clf, {})[stamp] = df
    return df


class _NoHutter:
    """view of a plain sklearn forest without Hutter tables"""

    def __init__(self, model):
        self.estimators_ = model.estimators_
        self.n_features_in_ = model.n_features_in_
        self.tree_means_per_leaf = None
        self.tree_vars_per_leaf = None


def rf_eval(forests, Xcm, want_leaves=False, want_mean=False, want_var=False, want_pred=False, acq=None, feas=None):
    """One launch of hm_rf_eval.  forests: list of DeviceForest; Xcm: float32 CUDA tensor [D_enc][N].
    Returns a dict of CUDA tensors."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    M = len(forests)
    assert Xcm.dtype == torch.float32 and Xcm.is_cuda and Xcm.dim() == 2 and Xcm.is_contiguous()
    D, N = Xcm.shape
    if D != forests[0].n_features:
        raise ValueError("X has %d features, but the forest is expecting %d features as input" % (D, forests[0].n_features))
    dev = Xcm.device
    out = {}
    handles = (ctypes.c_void_p * M)(*[f.handle for f in forests])
    leaf_ptrs = None
    if want_leaves:
        out["leaves"] = [torch.empty((f.n_trees, N), dtype=torch.int32, device=dev) for f in forests]
        leaf_ptrs = (ctypes.c_void_p * M)(*[t.data_ptr() for t in out["leaves"]])
    for key, flag in (("mean", want_mean), ("var", want_var), ("pred", want_pred)):
        if flag:
            out[key] = torch.empty((M, N), dtype=torch.float64, device=dev)
    acq_ref = None
    if acq is not None:
        out["score"] = torch.empty((N,), dtype=torch.float64, device=dev)
        acq_ref = ctypes.byref(acq)
    if N == 0:
        return out
    rc = L.hm_rf_eval(handles, M, _lib.dptr(Xcm), N, N, leaf_ptrs, _lib.dptr(out.get("mean")), _lib.dptr(out.get("var")),
                      _lib.dptr(out.get("pred")), acq_ref, _lib.dptr(feas), _lib.dptr(out.get("score")), _lib.stream_ptr())
    _lib.check(rc, "hm_rf_eval")
    return out


def attach_packing(forests, device_space):
    """hm_forest_attach_packing on every forest (idempotent): needed once before rf_eval_packed / score_packed"""
    L = _lib.lib()
    for f in forests:
        _lib.check(L.hm_forest_attach_packing(f.handle, device_space.handle), "hm_forest_attach_packing")
        f.packed_groups = int(L.hm_forest_packed_groups(f.handle))
        f.packed_kind = int(L.hm_forest_packed(f.handle))     # 1 FAST, 2 wide (64-bit shift), 3 FAST + self-looping leaves
    return forests


def rf_eval_packed(forests, codes, want_leaves=False, want_mean=False, want_var=False, want_pred=False, acq=None,
                   feas=None):
    """rf_eval on packed candidates: codes = CUDA int64 tensor [N] (64-bit codes of an all-discrete space,
    DeviceSpace.encode_packed / sample_pool_packed); the forests must have been through attach_packing."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    M = len(forests)
    assert codes.dtype == torch.int64 and codes.is_cuda and codes.dim() == 1 and codes.is_contiguous()
    N = codes.numel()
    dev = codes.device
    out = {}
    handles = (ctypes.c_void_p * M)(*[f.handle for f in forests])
    leaf_ptrs = None
    if want_leaves:
        out["leaves"] = [torch.empty((f.n_trees, N), dtype=torch.int32, device=dev) for f in forests]
        leaf_ptrs = (ctypes.c_void_p * M)(*[t.data_ptr() for t in out["leaves"]])
    for key, flag in (("mean", want_mean), ("var", want_var), ("pred", want_pred)):
        if flag:
            out[key] = torch.empty((M, N), dtype=torch.float64, device=dev)
    acq_ref = None
    if acq is not None:
        out["score"] = torch.empty((N,), dtype=torch.float64, device=dev)
        acq_ref = ctypes.byref(acq)
    if N == 0:
        return out
    rc = L.hm_rf_eval_packed(handles, M, _lib.dptr(codes), N, leaf_ptrs, _lib.dptr(out.get("mean")),
                             _lib.dptr(out.get("var")), _lib.dptr(out.get("pred")), acq_ref, _lib.dptr(feas),
                             _lib.dptr(out.get("score")), _lib.stream_ptr())
    _lib.check(rc, "hm_rf_eval_packed")
    return out


def topk(values, k, index_base=0):
    """Stable k smallest of a CUDA fp64 vector: (values[k], indices[k]) CUDA tensors, (+inf,-1) padded."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    assert values.dtype == torch.float64 and values.is_cuda and values.is_contiguous()
    N = values.numel()
    ws = torch.empty((int(L.hm_topk_workspace_bytes(k)),), dtype=torch.uint8, device=values.device)
    ov = torch.empty((k,), dtype=torch.float64, device=values.device)
    oi = torch.empty((k,), dtype=torch.int64, device=values.device)
    rc = L.hm_topk(_lib.dptr(values), N, int(index_base), int(k), _lib.dptr(ov), _lib.dptr(oi), _lib.dptr(ws), _lib.stream_ptr())
    _lib.check(rc, "hm_topk")
    return ov, oi


def topk_merge(values, indices, k):
    torch = _lib.require_cuda()
    L = _lib.lib()
    assert values.dtype == torch.float64 and indices.dtype == torch.int64 and values.is_cuda and indices.is_cuda
    values = values.contiguous().view(-1)
    indices = indices.contiguous().view(-1)
    ws = torch.empty((int(L.hm_topk_workspace_bytes(k)),), dtype=torch.uint8, device=values.device)
    ov = torch.empty((k,), dtype=torch.float64, device=values.device)
    oi = torch.empty((k,), dtype=torch.int64, device=values.device)
    rc = L.hm_topk_merge(_lib.dptr(values), _lib.dptr(indices), values.numel(), int(k), _lib.dptr(ov), _lib.dptr(oi),
                         _lib.dptr(ws), _lib.stream_ptr())
    _lib.check(rc, "hm_topk_merge")
    return ov, oi


_ST_WS = {}


def score_topk(forests, X, acq, k, feas=None):
    """hm_rf_score_topk / hm_rf_score_topk_packed: the k best candidates (values ascending, ties -> lowest index, NaN
    first: get_min_configurations, utility_functions.py:295-316) of a pool without materialising its score array.
    X: CUDA float32 [D_enc][N] (column-major encoded candidates) or CUDA int64 [N] (packed codes; the forests must have
    been through attach_packing).  -> (values fp64 [k], indices int64 [k]) on the device."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    M = len(forests)
    handles = (ctypes.c_void_p * M)(*[f.handle for f in forests])
    packed = X.dtype == torch.int64
    assert X.is_cuda and X.is_contiguous() and (X.dim() == 1 if packed else (X.dim() == 2 and X.dtype == torch.float32))
    N = X.shape[0] if packed else X.shape[1]
    dev = X.device
    key = (dev.index, int(k))
    ws = _ST_WS.get(key)
    if ws is None:
        ws = _ST_WS[key] = torch.empty((int(L.hm_rf_score_topk_workspace_bytes(int(k))),), dtype=torch.uint8, device=dev)
    vals = torch.empty((k,), dtype=torch.float64, device=dev)
    idx = torch.empty((k,), dtype=torch.int64, device=dev)
    if packed:
        rc = L.hm_rf_score_topk_packed(handles, M, _lib.dptr(X), N, ctypes.byref(acq), _lib.dptr(feas), int(k),
                                       _lib.dptr(vals), _lib.dptr(idx), _lib.dptr(ws), _lib.stream_ptr())
    else:
        rc = L.hm_rf_score_topk(handles, M, _lib.dptr(X), N, N, ctypes.byref(acq), _lib.dptr(feas), int(k),
                                _lib.dptr(vals), _lib.dptr(idx), _lib.dptr(ws), _lib.stream_ptr())
    _lib.check(rc, "hm_rf_score_topk")
    return vals, idx
