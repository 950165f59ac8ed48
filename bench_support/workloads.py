"""Synthetic workloads of the shapes BASELINE.json names: scenario JSON, seeded training
data, fitted surrogate, seeded candidate pools, acquisition parameters.  Used by bench.py, smoke() and the
large-size tests; everything here is host-side setup outside any timed region."""
import math

import numpy as np

from hypermapper_b200 import _lib
from hypermapper_b200.encoding import Encoder

from .hutter_forest import fit_hutter_forest
from .space_lite import SpaceLite


def _ordinal(values):
    return {"parameter_type": "ordinal", "values": list(values), "parameter_default": values[0]}


def _categorical(n, prefix):
    vals = ["%s%d" % (prefix, i) for i in range(n)]
    return {"parameter_type": "categorical", "values": vals, "parameter_default": vals[0]}


def scenario_c1():
    """hypermapper/_branin_scenario.json shape: 2 real inputs, RF (10 trees), 1 objective, EI"""
    return {
        "application_name": "c1_branin", "optimization_objectives": ["Value"],
        "models": {"model": "random_forest", "number_of_trees": 10},
        "input_parameters": {
            "x1": {"parameter_type": "real", "values": [-5, 10], "parameter_default": 0},
            "x2": {"parameter_type": "real", "values": [0, 15], "parameter_default": 0}},
    }


def scenario_c3():
    """2-objective mixed ordinal/categorical space: 6 ordinal (4,8,16,16,32,64 values) + 4 categorical (2,3,4,8)
    -> D_enc = 6 + 17 = 23; 256-tree forests."""
    ip = {}
    for j, size in enumerate((4, 8, 16, 16, 32, 64)):
        ip["o%d" % j] = _ordinal([2 ** i for i in range(size)] if size <= 16 else list(range(1, size + 1)))
    import os
    cats = tuple(int(v) for v in os.environ.get("HM_BENCH_C3_CATS", "2,3,4,8").split(","))      # experiments only
    for j, size in enumerate(cats):
        ip["c%d" % j] = _categorical(size, "v")
    return {"application_name": "c3_mixed", "optimization_objectives": ["obj0", "obj1"],
            "models": {"model": "random_forest", "number_of_trees": 256}, "input_parameters": ip}


def scenario_c5():
    """20-D mixed space: 8 real [0,1], 4 integer [0,31], 4 ordinal (16 values), 4 categorical (4 values)
    -> D_enc = 16 + 16 = 32"""
    ip = {}
    for j in range(8):
        ip["r%d" % j] = {"parameter_type": "real", "values": [0, 1], "parameter_default": 0.5}
    for j in range(4):
        ip["i%d" % j] = {"parameter_type": "integer", "values": [0, 31], "parameter_default": 0}
    for j in range(4):
        ip["o%d" % j] = _ordinal(list(range(1, 17)))
    for j in range(4):
        ip["c%d" % j] = _categorical(4, "v")
    return {"application_name": "c5_mixed20", "optimization_objectives": ["obj0"],
            "models": {"model": "random_forest", "number_of_trees": 256}, "input_parameters": ip}


def scenario_c2():
    """6-D Hartmann on [0,1]^6, GP surrogate"""
    ip = {"x%d" % j: {"parameter_type": "real", "values": [0, 1], "parameter_default": 0.5} for j in range(6)}
    return {"application_name": "c2_hartmann6", "optimization_objectives": ["Value"],
            "models": {"model": "gaussian_process"}, "input_parameters": ip}


def hartmann6(X):
    alpha = np.array([1.0, 1.2, 3.0, 3.2])
    A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]])
    P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                         [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
    X = np.atleast_2d(X)
    inner = np.einsum("ij,nij->ni", A, (X[:, None, :] - P[None]) ** 2)
    return -np.sum(alpha * np.exp(-inner), axis=1)


def raw_uniform_columns(space, n, rng):
    """{input name: array} drawn uniformly per parameter (index draws for discrete types)."""
    cols = {}
    objs = space.get_input_parameters_objects()
    for p in space.get_input_parameters():
        t = space.get_type(p)
        o = objs[p]
        if t == "real":
            cols[p] = rng.random(n) * (o.get_max() - o.get_min()) + o.get_min()
        elif t == "integer":
            cols[p] = rng.integers(o.get_min(), o.get_max() + 1, n)
        elif t == "ordinal":
            cols[p] = np.asarray(o.get_values())[rng.integers(0, o.get_size(), n)]
        else:
            cols[p] = rng.integers(0, o.get_size(), n)
    return cols


def synthetic_objectives(space, enc_rows, n_obj, seed):
    """seeded weighted sums + pairwise interactions of the encoded features, one column per objective"""
    rng = np.random.default_rng(seed)
    D = enc_rows.shape[1]
    out = []
    for _ in range(n_obj):
        w = rng.standard_normal(D)
        pairs = rng.integers(0, D, (6, 2))
        pw = rng.standard_normal(6) * 2
        y = enc_rows @ w
        for (a, b), c in zip(pairs, pw):
            y = y + c * enc_rows[:, a] * enc_rows[:, b]
        out.append(y + 0.05 * rng.standard_normal(len(y)))
    return out


class RFWorkload:
    """Fitted Hutter forests + acquisition parameters for one of the RF scenarios."""

    def __init__(self, scenario, n_train, seed, acquisition="EI", scalarization="tchebyshev", iteration_number=5):
        self.config = scenario
        self.space = SpaceLite(scenario)
        self.encoder = Encoder(self.space)
        self.objectives = list(scenario["optimization_objectives"])
        self.D_enc = self.encoder.n_encoded
        self.n_trees = scenario["models"]["number_of_trees"]
        rng = np.random.default_rng(seed)
        cols = raw_uniform_columns(self.space, n_train, rng)
        Xtr = np.ascontiguousarray(self.encoder.encode_columns(cols).T)
        if scenario["application_name"] == "c1_branin":
            x1, x2 = cols["x1"], cols["x2"]
            a, b, c, r, s, t = 1.0, 5.1 / (4 * math.pi ** 2), 5 / math.pi, 6.0, 10.0, 1 / (8 * math.pi)
            ys = [a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s]
        else:
            ys = synthetic_objectives(self.space, Xtr, len(self.objectives), seed + 17)
        self.train_cols = cols
        self.data_array = {o: [float(v) for v in y] for o, y in zip(self.objectives, ys)}
        self.models = {}
        for j, o in enumerate(self.objectives):
            self.models[o] = fit_hutter_forest(Xtr, ys[j], n_estimators=self.n_trees, max_features=0.5,
                                               min_samples_split=5, bootstrap=False, seed=seed + 31 * j)
        # weights: one Dirichlet(1,..,1) draw (utility_functions.sample_weight_flat, utility_functions.py:804-824)
        w = np.random.default_rng(seed + 5).dirichlet(np.ones(len(self.objectives)))
        self.objective_weights = {o: float(x) for o, x in zip(self.objectives, w)}
        self.objective_limits = {o: [min(self.data_array[o]), max(self.data_array[o])] for o in self.objectives}
        self.acquisition = acquisition
        self.scalarization = scalarization
        self.iteration_number = iteration_number

    def candidates_raw(self, n, seed, out=None):
        """float64 [n][D] raw parameter values of the SAME candidates candidates_encoded(n, seed) encodes"""
        rng = np.random.default_rng(seed)
        inputs = self.space.get_input_parameters()
        if out is None:
            out = np.empty((n, len(inputs)), dtype=np.float64)
        chunk = 1 << 21
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            cols = raw_uniform_columns(self.space, hi - lo, rng)
            for j, p in enumerate(inputs):
                out[lo:hi, j] = cols[p]
        return out

    def candidates_encoded(self, n, seed, out=None):
        """float32 [D_enc][n] column-major candidates, uniform over every parameter's values."""
        rng = np.random.default_rng(seed)
        if out is None:
            out = np.empty((self.D_enc, n), dtype=np.float32)
        chunk = 1 << 21
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            cols = raw_uniform_columns(self.space, hi - lo, rng)
            out[:, lo:hi] = self.encoder.encode_columns(cols, dtype=np.float32)
        return out

    def acq_params(self):
        from hypermapper_b200.random_scalarizations import acquisition_constants
        w, fmin, beta = acquisition_constants(self.acquisition, self.scalarization, self.objectives,
                                              self.objective_weights, self.objective_limits, self.data_array,
                                              self.iteration_number)
        return _lib.make_acq_params(self.acquisition, self.scalarization, w, fmin, beta)
