"""Input encoding: the host side of models.preprocess_data_array / preprocess_data_buffer (models.py:921-984).

Encoded column order (models.py:957-984): every non-categorical input in input-parameter order
(ordinal -> values.index(v)/size evaluated in float64; real/integer -> raw value, or the z-score when
normalize_inputs is set), THEN one one-hot block per categorical input (value == integer index).

The output is column-major ([D_enc][N]) because that is the layout the kernels read: a warp reads 32
consecutive candidates of one feature in one 128-byte transaction.
"""
import numpy as np


class Encoder:
    def __init__(self, param_space):
        self.space = param_space
        self.inputs = list(param_space.get_input_parameters())
        objs = param_space.get_input_parameters_objects()
        self.col_of = {p: j for j, p in enumerate(self.inputs)}
        self.non_cat = list(param_space.get_input_non_categorical_parameters(self.inputs))
        self.cat = list(param_space.get_input_categorical_parameters(self.inputs))
        self.kinds = {p: param_space.get_type(p) for p in self.inputs}
        self.ord_values = {}
        self.ord_lut = {}
        for p in self.non_cat:
            if self.kinds[p] == "ordinal":
                vals = list(objs[p].get_values())
                size = objs[p].get_size()
                self.ord_values[p] = vals
                # index/size in Python float64 arithmetic, exactly as models.py:962-965
                self.ord_lut[p] = np.array([i / size for i in range(len(vals))], dtype=np.float64)
        self.cat_sizes = {p: objs[p].get_size() for p in self.cat}
        self.n_encoded = len(self.non_cat) + sum(self.cat_sizes.values())
        self.names = list(self.non_cat)
        for p in self.cat:
            self.names += ["%s_%d" % (p, c) for c in range(self.cat_sizes[p])]

    # -- column access -------------------------------------------------------------------------
    def _columns_from_bufferx(self, bufferx):
        """bufferx: list of tuples/lists in input-parameter order, or a 2-D array [N][D]."""
        if isinstance(bufferx, np.ndarray) and bufferx.ndim == 2:
            return {p: bufferx[:, j] for p, j in self.col_of.items()}
        n = len(bufferx)
        cols = {p: [None] * n for p in self.inputs}
        for r, entry in enumerate(bufferx):
            for j, p in enumerate(self.inputs):
                cols[p][r] = entry[j]
        return cols

    def _ordinal_index(self, p, col):
        vals = self.ord_values[p]
        arr = np.asarray(col)
        if arr.dtype == object:
            arr = arr.astype(np.float64)
        sv = np.asarray(vals, dtype=np.float64)
        idx = np.searchsorted(sv, arr.astype(np.float64), side="left")
        idx = np.clip(idx, 0, len(sv) - 1)
        if not np.array_equal(sv[idx], arr.astype(np.float64)):
            bad = arr[sv[idx] != arr.astype(np.float64)][0]
            raise ValueError("%r is not in list" % (bad,))   # what list.index raises in the reference
        return idx

    def encode_columns(self, cols, dtype=np.float64):
        """cols: {input name: sequence}. Returns a C-contiguous [D_enc][N] array of `dtype`."""
        first = cols[self.inputs[0]]
        n = len(first)
        out = np.empty((self.n_encoded, n), dtype=dtype)
        j = 0
        norm = self.space.get_input_normalization_flag() is True
        for p in self.non_cat:
            if self.kinds[p] == "ordinal":
                out[j] = self.ord_lut[p][self._ordinal_index(p, cols[p])]
            elif norm:
                x = np.array(cols[p], dtype=np.float64)
                out[j] = (x - self.space.get_parameter_mean(p)) / self.space.get_parameter_std(p)
            else:
                c = cols[p]
                out[j] = c.astype(np.float64) if isinstance(c, np.ndarray) else np.array(c, dtype=np.float64)
            j += 1
        for p in self.cat:
            x = np.asarray(cols[p])
            if x.dtype == object or x.dtype.kind == "f":
                x = x.astype(np.int64)
            size = self.cat_sizes[p]
            if n and (x.min() < 0 or x.max() >= size):
                # sklearn's OneHotEncoder raises on unknown categories (models.py:977-980)
                raise ValueError("Found unknown categories in column %s during transform" % p)
            for c in range(size):
                out[j] = (x == c)
                j += 1
        return out

    def encode(self, bufferx, dtype=np.float64):
        return self.encode_columns(self._columns_from_bufferx(bufferx), dtype=dtype)

    def encode_configurations(self, configurations, dtype=np.float64):
        """configurations: list of dicts (the local_search plug-in convention, local_search.py:468-473)"""
        n = len(configurations)
        cols = {p: [None] * n for p in self.inputs}
        for r, c in enumerate(configurations):
            for p in self.inputs:
                cols[p][r] = c[p]
        return self.encode_columns(cols, dtype=dtype)


def raw_matrix(bufferx, param_space):
    """list of dicts / list of tuples / [N][D] array -> [N][D] float64 of raw values in input-parameter order
    (categoricals as integer indices, space.py:669-673)."""
    inputs = list(param_space.get_input_parameters())
    if isinstance(bufferx, np.ndarray):
        arr = bufferx
    elif len(bufferx) > 0 and isinstance(bufferx[0], dict):
        arr = np.array([[c[p] for p in inputs] for c in bufferx])
    else:
        arr = np.asarray(bufferx)
    if arr.dtype != np.float64:
        arr = arr.astype(np.float64)
    return arr.reshape(-1, len(inputs))


_encoders = {}


def encoder_for(param_space):
    """One Encoder per Space object (the LUTs depend only on the space definition, and on the
    normalisation statistics which are read at encode time)."""
    key = id(param_space)
    enc = _encoders.get(key)
    if enc is None or enc.space is not param_space:
        enc = Encoder(param_space)
        _encoders[key] = enc
    return enc
