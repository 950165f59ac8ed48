"""The search-space side of the hot path on the device (csrc/hm_space.cu):

  encode        models.preprocess_data_array (models.py:941-984)                            K5 hm_encode
  sample_pool   space.get_random_configuration(size, use_priors, return_as_array=True)
                (space.py:1263-1311) + the encoding, in one pass                            K6 hm_sample_pool
  prior_prob    prior_optimization.compute_probability_from_prior (:65-95)                  K7 hm_prior_prob

A `DeviceSpace` is built from any object with the reference `Space` accessor surface (the real `space.Space`
or `space_lite.SpaceLite`): the descriptors hold only numbers, every table is computed on the host BY THE
PARAMETER OBJECTS THEMSELVES (ordinal index/size, get_x_probability per discrete value), so the device
arithmetic that remains is the z-score, the Beta/Gaussian/interpolated densities of real parameters and the
product over parameters.
"""
import ctypes
import math

import numpy as np

from . import _lib

P_REAL, P_INTEGER, P_ORDINAL, P_CATEGORICAL = 0, 1, 2, 3
PRIOR_BETA, PRIOR_TABLE, PRIOR_GAUSSIAN, PRIOR_INTERP = 0, 1, 2, 3
KIND_OF = {"real": P_REAL, "integer": P_INTEGER, "ordinal": P_ORDINAL, "categorical": P_CATEGORICAL}
# integer parameters with a named density get a per-value probability table up to this many values
INTEGER_TABLE_LIMIT = 1 << 16


class ParamDesc(ctypes.Structure):
    """hm_param_desc_t (include/hm_b200.h)"""
    _fields_ = [("kind", ctypes.c_int32), ("n_values", ctypes.c_int32), ("out_col", ctypes.c_int32),
                ("normalize", ctypes.c_int32), ("prior_kind", ctypes.c_int32), ("n_prior", ctypes.c_int32),
                ("ordinal_beta", ctypes.c_int32), ("n_cdf", ctypes.c_int32), ("values_off", ctypes.c_int64),
                ("prior_off", ctypes.c_int64), ("cdf_off", ctypes.c_int64), ("lo", ctypes.c_double),
                ("hi", ctypes.c_double), ("mean", ctypes.c_double), ("std", ctypes.c_double),
                ("a", ctypes.c_double), ("b", ctypes.c_double), ("lbeta", ctypes.c_double)]


def _lbeta(a, b):
    return math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)


def _cumulative(p):
    """the table np.random.choice searches (numpy/random/mtrand.pyx: cdf = p.cumsum(); cdf /= cdf[-1])"""
    cdf = np.asarray(p, dtype=np.float64).cumsum()
    return cdf / cdf[-1]


def describe_space(param_space):
    """-> (ctypes array of ParamDesc in input-parameter order, fp64 table blob, n_encoded).
    Raises HmError for priors that have no device form (multivariate KDE 'estimate', Gaussian mixtures)."""
    inputs = list(param_space.get_input_parameters())
    objs = param_space.get_input_parameters_objects()
    non_cat = list(param_space.get_input_non_categorical_parameters(inputs))
    cat = list(param_space.get_input_categorical_parameters(inputs))
    if param_space.get_estimate_prior_flags()[0] if hasattr(param_space, "get_estimate_prior_flags") else False:
        raise _lib.HmError("estimated (multivariate KDE) priors are evaluated by the reference's space.pdf; "
                           "they have no device form")
    col = {}
    j = 0
    for p in non_cat:
        col[p] = j
        j += 1
    for p in cat:
        col[p] = j
        j += objs[p].get_size()
    n_encoded = j
    norm = param_space.get_input_normalization_flag() is True
    descs = (ParamDesc * len(inputs))()
    tables = []
    n_tab = 0

    def push(arr):
        nonlocal n_tab
        arr = np.asarray(arr, dtype=np.float64).ravel()
        off = n_tab
        tables.append(arr)
        n_tab += arr.size
        return off

    for i, name in enumerate(inputs):
        obj = objs[name]
        kind = KIND_OF[param_space.get_type(name)]
        d = descs[i]
        d.kind = kind
        d.out_col = col[name]
        d.a = d.b = 1.0
        prior = getattr(obj, "prior", "uniform")
        if kind in (P_REAL, P_INTEGER):
            d.lo, d.hi = float(obj.get_min()), float(obj.get_max())
            if norm:
                d.normalize = 1
                d.mean, d.std = float(param_space.get_parameter_mean(name)), float(param_space.get_parameter_std(name))
        if kind == P_REAL:
            if isinstance(prior, list):
                d.prior_kind = PRIOR_INTERP
                d.n_prior = len(prior)
                d.prior_off = push(np.concatenate([np.linspace(obj.get_min(), obj.get_max(), len(prior)),
                                                   np.asarray(prior, dtype=np.float64)]))
                d.n_cdf = len(obj.param_xs)
                d.cdf_off = push(np.concatenate([obj.cdf_distribution, obj.param_xs]))
            elif prior == "custom_gaussian":
                mean, std = obj.custom_gaussian_prior_mean, obj.custom_gaussian_prior_std
                if isinstance(mean, (list, tuple)) or isinstance(std, (list, tuple)):
                    raise _lib.HmError("Gaussian-mixture priors have no device form")
                d.prior_kind = PRIOR_GAUSSIAN
                d.a, d.b = float(mean), float(std)
            elif prior == "estimate":
                raise _lib.HmError("estimated (KDE) priors have no device form")
            else:
                d.prior_kind = PRIOR_BETA
                d.a, d.b = float(obj.alpha), float(obj.beta)
        elif kind == P_INTEGER:
            size = int(obj.get_max()) - int(obj.get_min()) + 1
            if size > 2 ** 31 - 1:
                # hm_sample_pool draws integer values with one 32-bit Philox word (csrc/hm_space.cu)
                raise _lib.HmError("integer parameter %r spans %d values; the device sampler supports at most 2^31 - 1"
                                   % (name, size))
            if prior == "distribution":
                d.prior_kind = PRIOR_TABLE
                d.n_prior = size
                probs = np.asarray(obj.distribution, dtype=np.float64)
                d.prior_off = push(np.concatenate([probs, _cumulative(probs)]))
            else:
                d.a, d.b = float(obj.alpha), float(obj.beta)
                if size <= INTEGER_TABLE_LIMIT:
                    d.prior_kind = PRIOR_TABLE
                    d.ordinal_beta = 1
                    d.n_prior = size
                    probs = np.array([obj.get_x_probability(v) for v in range(int(obj.get_min()), int(obj.get_max()) + 1)],
                                     dtype=np.float64)
                    d.prior_off = push(np.concatenate([probs, np.zeros(size)]))
                else:
                    d.prior_kind = PRIOR_BETA
        else:
            values = list(obj.get_values()) if kind == P_ORDINAL else list(range(obj.get_size()))
            d.n_values = len(values)
            if kind == P_ORDINAL:
                size = obj.get_size()
                d.values_off = push(np.concatenate([np.asarray(values, dtype=np.float64),
                                                    np.array([k / size for k in range(len(values))])]))
                if len(values) > 1 and not np.all(np.diff(np.asarray(values, dtype=np.float64)) > 0):
                    raise _lib.HmError("ordinal values of %s are not strictly ascending" % name)
            d.prior_kind = PRIOR_TABLE
            d.n_prior = len(values)
            probs = np.array([obj.get_x_probability(v) for v in values], dtype=np.float64)
            named = kind == P_ORDINAL and prior != "distribution"
            if named:
                d.ordinal_beta = 1
                d.a, d.b = float(obj.densities_alphas[prior]), float(obj.densities_betas[prior])
                d.prior_off = push(np.concatenate([probs, np.zeros(len(values))]))
            else:
                d.prior_off = push(np.concatenate([probs, _cumulative(probs)]))
        d.lbeta = _lbeta(d.a, d.b) if d.prior_kind == PRIOR_BETA else 0.0
    blob = np.concatenate(tables) if tables else np.zeros(0)
    return descs, np.ascontiguousarray(blob, dtype=np.float64), n_encoded


class DeviceSpace:
    def __init__(self, param_space):
        L = _lib.lib()
        _lib.require_cuda()
        self.space = param_space
        self.inputs = list(param_space.get_input_parameters())
        self.descs, self.tables, self.n_encoded = describe_space(param_space)
        self.D = len(self.inputs)
        self.stats_key = _stats_key(param_space)
        h = ctypes.c_void_p()
        rc = L.hm_space_create(self.D, self.descs, self.tables.ctypes.data_as(ctypes.c_void_p), self.tables.size,
                               ctypes.byref(h))
        _lib.check(rc, "hm_space_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().hm_space_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- helpers --------------------------------------------------------------------------------------
    def raw_to_device(self, raw):
        """raw: [N][D] array-like of raw values (categoricals as indices) -> CUDA fp64 tensor, row-major."""
        torch = _lib.require_cuda()
        if torch.is_tensor(raw):
            return raw.to(device="cuda", dtype=torch.float64).contiguous()
        arr = np.asarray(raw)
        if arr.dtype != np.float64:
            arr = arr.astype(np.float64)
        if arr.ndim != 2 or arr.shape[1] != self.D:
            if arr.size == 0:
                arr = arr.reshape(0, self.D)
            else:
                raise ValueError("expected [N][%d] raw parameter values, got shape %r" % (self.D, arr.shape))
        return torch.from_numpy(np.ascontiguousarray(arr)).cuda()

    def _invalid_counter(self):
        torch = _lib.require_cuda()
        return torch.zeros((1,), dtype=torch.int32, device="cuda")

    # -- K5 -------------------------------------------------------------------------------------------
    def encode(self, raw, want32=True, want64=False, col_major=False):
        """-> (x32 [D_enc][N] or None, x64 [D_enc][N] or None).  Raises ValueError like the reference
        (list.index / OneHotEncoder) when a discrete value is not one of the parameter's values.
        col_major: raw is a CUDA fp64 tensor [D][N] (the layout of device pools)."""
        torch = _lib.require_cuda()
        rd = raw if col_major else self.raw_to_device(raw)
        N = rd.shape[1] if col_major else rd.shape[0]
        x32 = torch.empty((self.n_encoded, N), dtype=torch.float32, device="cuda") if want32 else None
        x64 = torch.empty((self.n_encoded, N), dtype=torch.float64, device="cuda") if want64 else None
        if N == 0:
            return x32, x64
        bad = self._invalid_counter()
        rc = _lib.lib().hm_encode(self.handle, _lib.dptr(rd), 1 if col_major else 0, N if col_major else self.D, N,
                                  _lib.dptr(x32), N, _lib.dptr(x64), N, _lib.dptr(bad), _lib.stream_ptr())
        _lib.check(rc, "hm_encode")
        if int(bad.item()) != 0:
            raise ValueError("a discrete parameter value is not in its list of values (models.py:962-983)")
        return x32, x64

    # -- packed candidates (all-discrete spaces; include/hm_b200.h "Packed candidates", csrc/hm_pack.cuh) -------
    @property
    def packed_bits(self):
        """bits of a candidate's 64-bit code, 0 when the space cannot be packed (real / integer parameters, > 64 bits)"""
        return int(_lib.lib().hm_space_packed_bits(self.handle))

    def packed_fields(self):
        """[(low bit, width)] per input parameter"""
        out = []
        for j in range(self.D):
            lo, w = ctypes.c_int(0), ctypes.c_int(0)
            _lib.check(_lib.lib().hm_space_packed_field(self.handle, j, ctypes.byref(lo), ctypes.byref(w)),
                       "hm_space_packed_field")
            out.append((lo.value, w.value))
        return out

    def encode_packed(self, raw):
        """raw [N][D] values (host or device) -> CUDA uint64-as-int64 tensor [N] of packed codes (hm_encode_packed)."""
        torch = _lib.require_cuda()
        rd = self.raw_to_device(raw)
        N = rd.shape[0]
        codes = torch.empty((N,), dtype=torch.int64, device="cuda")
        if N == 0:
            return codes
        bad = self._invalid_counter()
        rc = _lib.lib().hm_encode_packed(self.handle, _lib.dptr(rd), 0, self.D, N, _lib.dptr(codes), _lib.dptr(bad),
                                         _lib.stream_ptr())
        _lib.check(rc, "hm_encode_packed")
        if int(bad.item()) != 0:
            raise ValueError("a discrete parameter value is not in its list of values (models.py:962-983)")
        return codes

    def sample_pool_packed(self, seed, N, use_priors, index_base=0, want_raw=False):
        """hm_sample_pool_packed: the same pool as sample_pool (same Philox stream), as packed codes.
        -> dict(codes=int64 [N], raw=[D][N] fp64 | None)"""
        torch = _lib.require_cuda()
        codes = torch.empty((N,), dtype=torch.int64, device="cuda")
        raw = torch.empty((self.D, N), dtype=torch.float64, device="cuda") if want_raw else None
        rc = _lib.lib().hm_sample_pool_packed(self.handle, ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), int(index_base), N,
                                              1 if use_priors else 0, _lib.dptr(codes), _lib.dptr(raw), N,
                                              _lib.stream_ptr())
        _lib.check(rc, "hm_sample_pool_packed")
        return {"codes": codes, "raw": raw}

    def pack_codes_host(self, raw, out=None):
        """numpy restatement of the packing for host-resident candidates: raw [N][D] values (categoricals as indices,
        ordinals as values) -> uint64 [N].  What a caller that samples on the host hands to hm_rf_score_packed_host."""
        raw = np.asarray(raw, dtype=np.float64)
        N = raw.shape[0]
        codes = np.zeros((N,), dtype=np.uint64) if out is None else out
        if out is not None:
            codes[:] = 0
        objs = self.space.get_input_parameters_objects()
        for j, ((low, width), name) in enumerate(zip(self.packed_fields(), self.inputs)):
            col = raw[:, j]
            if self.descs[j].kind == P_ORDINAL:
                sv = np.asarray(objs[name].get_values(), dtype=np.float64)
                idx = np.clip(np.searchsorted(sv, col), 0, len(sv) - 1)
                if not np.array_equal(sv[idx], col):
                    raise ValueError("a discrete parameter value is not in its list of values (models.py:962-983)")
                codes |= idx.astype(np.uint64) << np.uint64(low)
            else:
                idx = col.astype(np.int64)
                n_cat = int(self.descs[j].n_values)
                if N and (idx.min() < 0 or idx.max() >= n_cat or not np.array_equal(idx, col)):
                    raise ValueError("a categorical index is out of range")
                # one-hot over the first n - 1 categories; the last category is the all-zero field (csrc/hm_pack.cuh)
                bit = np.where(idx < n_cat - 1, np.left_shift(np.uint64(1), np.minimum(idx, 62).astype(np.uint64)),
                               np.uint64(0)).astype(np.uint64)
                codes |= bit << np.uint64(low)
        return codes

    def unpack_codes_host(self, codes):
        """inverse of pack_codes_host: uint64 [n] -> raw fp64 rows [n][D]"""
        codes = np.asarray(codes).astype(np.uint64)
        objs = self.space.get_input_parameters_objects()
        out = np.empty((len(codes), self.D), dtype=np.float64)
        for j, ((low, width), name) in enumerate(zip(self.packed_fields(), self.inputs)):
            field = (codes >> np.uint64(low)) & np.uint64((1 << width) - 1)
            if self.descs[j].kind == P_ORDINAL:
                out[:, j] = np.asarray(objs[name].get_values(), dtype=np.float64)[field.astype(np.int64)]
            else:
                n_cat = int(self.descs[j].n_values)
                out[:, j] = np.where(field == 0, float(n_cat - 1), np.log2(np.maximum(field, 1).astype(np.float64)))
        return out

    # -- K7 -------------------------------------------------------------------------------------------
    def prior_prob(self, raw, objective_weights, col_major=False):
        """compute_probability_from_prior on the device -> CUDA fp64 tensor [N].
        raw: [N][D] values (host or device), or with col_major a CUDA fp64 tensor [D][N]."""
        torch = _lib.require_cuda()
        rd = raw if col_major else self.raw_to_device(raw)
        N = rd.shape[1] if col_major else rd.shape[0]
        out = torch.empty((N,), dtype=torch.float64, device="cuda")
        if N == 0:
            return out
        w = (ctypes.c_double * len(objective_weights))(*[float(x) for x in objective_weights])
        bad = self._invalid_counter()
        rc = _lib.lib().hm_prior_prob(self.handle, _lib.dptr(rd), 1 if col_major else 0, N if col_major else self.D, N,
                                      w, len(objective_weights), _lib.dptr(out), _lib.dptr(bad), _lib.stream_ptr())
        _lib.check(rc, "hm_prior_prob")
        if int(bad.item()) != 0:
            raise ValueError("a discrete parameter value is not in its list of values")
        return out

    # -- K6 -------------------------------------------------------------------------------------------
    def sample_pool(self, seed, N, use_priors, index_base=0, want_raw=False, want32=True, want64=False,
                    prob_weights=None):
        """Pool candidates index_base .. index_base+N-1 of stream (seed, use_priors).
        -> dict(raw=[D][N] fp64 | None, x32, x64, prob)"""
        torch = _lib.require_cuda()
        raw = torch.empty((self.D, N), dtype=torch.float64, device="cuda") if want_raw else None
        x32 = torch.empty((self.n_encoded, N), dtype=torch.float32, device="cuda") if want32 else None
        x64 = torch.empty((self.n_encoded, N), dtype=torch.float64, device="cuda") if want64 else None
        prob = torch.empty((N,), dtype=torch.float64, device="cuda") if prob_weights is not None else None
        w, nw = None, 0
        if prob_weights is not None:
            nw = len(prob_weights)
            w = (ctypes.c_double * nw)(*[float(x) for x in prob_weights])
        rc = _lib.lib().hm_sample_pool(self.handle, ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), int(index_base), N,
                                       1 if use_priors else 0, _lib.dptr(raw), N, _lib.dptr(x32), N, _lib.dptr(x64),
                                       N, w, nw, _lib.dptr(prob), _lib.stream_ptr())
        _lib.check(rc, "hm_sample_pool")
        return {"raw": raw, "x32": x32, "x64": x64, "prob": prob}

    def sample_rows(self, seed, use_priors, indices):
        """Re-generate the raw rows of the given pool indices -> numpy [n][D] fp64."""
        torch = _lib.require_cuda()
        idx = torch.as_tensor(indices, dtype=torch.int64, device="cuda").contiguous()
        n = idx.numel()
        out = torch.empty((n, self.D), dtype=torch.float64, device="cuda")
        rc = _lib.lib().hm_sample_rows(self.handle, ctypes.c_uint64(int(seed) & (2 ** 64 - 1)),
                                       1 if use_priors else 0, _lib.dptr(idx), n, _lib.dptr(out), _lib.stream_ptr())
        _lib.check(rc, "hm_sample_rows")
        return out.cpu().numpy()

    def rows_to_configurations(self, rows):
        """raw fp64 rows -> list of dicts with the python types the reference uses (space.py:669-673)."""
        out = []
        kinds = [self.descs[j].kind for j in range(self.D)]
        objs = self.space.get_input_parameters_objects()
        for r in rows:
            c = {}
            for j, name in enumerate(self.inputs):
                v = float(r[j])
                if kinds[j] == P_REAL:
                    c[name] = v
                elif kinds[j] == P_ORDINAL:
                    vals = objs[name].get_values()
                    c[name] = vals[int(np.searchsorted(np.asarray(vals, dtype=np.float64), v))]
                else:
                    c[name] = int(v)
            out.append(c)
        return out


class DevicePool:
    """One random pool of local_search that lives on the GPU only: generated and encoded by hm_sample_pool, scored in
    place; single rows are re-generated from (seed, index) when the optimiser needs them as configurations."""

    def __init__(self, device_space, seed, use_priors, n, want64=False):
        from .models import EncodedCandidates
        self.ds, self.seed, self.use_priors, self.n = device_space, int(seed), bool(use_priors), int(n)
        if device_space.packed_bits > 0 and not want64:
            # all-discrete space: the pool is one 64-bit code per candidate (same Philox stream, same candidates)
            out = device_space.sample_pool_packed(seed, n, use_priors, want_raw=True)
            self.candidates = EncodedCandidates(raw_cm=out["raw"], codes=out["codes"], device_space=device_space)
        else:
            out = device_space.sample_pool(seed, n, use_priors, want_raw=True, want32=True, want64=want64)
            self.candidates = EncodedCandidates(x32=out["x32"], x64=out["x64"], raw_cm=out["raw"],
                                                device_space=device_space)

    def __len__(self):
        return self.n

    def rows(self, indices):
        """-> list of configuration dicts of the given pool rows"""
        idx = np.asarray(list(indices), dtype=np.int64)
        if idx.size == 0:
            return []
        return self.ds.rows_to_configurations(self.ds.sample_rows(self.seed, self.use_priors, idx))


def _stats_key(param_space):
    if param_space.get_input_normalization_flag() is not True:
        return None
    names = param_space.get_input_non_categorical_parameters(param_space.get_input_parameters())
    out = []
    for p in names:
        if param_space.get_type(p) in ("real", "integer"):
            out.append((param_space.get_parameter_mean(p), param_space.get_parameter_std(p)))
    return tuple(out)


_spaces = {}


def device_space_for(param_space):
    """One DeviceSpace per Space object; rebuilt when the z-score statistics change (models.py:966-971 reads
    them at encode time)."""
    key = id(param_space)
    ds = _spaces.get(key)
    if ds is None or ds.space is not param_space or ds.stats_key != _stats_key(param_space):
        ds = DeviceSpace(param_space)
        _spaces[key] = ds
    return ds
