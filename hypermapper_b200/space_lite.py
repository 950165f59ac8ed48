"""A small stand-in for the reference's `space.Space` exposing only the accessor surface the hot path uses.

The search-space model (space.py, 2774 lines: priors, KDEs, CSV IO, black-box dispatch ...) is OUT OF SCOPE and
stays with the reference; in a deployment the real `Space` object is passed in and every function of this
package duck-types against it.  `SpaceLite` exists so that tests, smoke() and bench.py can run on a box where
the reference is not installed.  It is built from the same scenario JSON ("input_parameters",
"optimization_objectives", "feasible_output", "normalize_inputs"; schema.json is unchanged) and reproduces the
accessor semantics cited below; priors other than uniform are not modelled.
"""
import itertools
from collections import OrderedDict, defaultdict

import numpy as np


class _Param:
    kind = None

    def get_default(self):
        return self.default


class RealParam(_Param):  # space.py:88-343
    kind = "real"

    def __init__(self, lo, hi, default=None):
        self.min_value, self.max_value, self.default = lo, hi, default
        self.preferred_discretization = np.linspace(lo, hi, num=10)

    def get_size(self):
        return float("inf")

    def get_discrete_size(self):
        return len(self.preferred_discretization)

    def get_discrete_values(self):
        return self.preferred_discretization

    def get_min(self):
        return self.min_value

    def get_max(self):
        return self.max_value

    def from_range_0_1_to_parameter_value(self, X):   # space.py:316-331
        Xc = np.array([X]).copy() if type(X) != np.ndarray and type(X) != list else np.array(X).copy()
        s = Xc * (self.max_value - self.min_value) + self.min_value
        return s[0] if len(Xc) == 1 else s

    def randomly_select_uniform(self, size=1):        # space.py:201-209
        return self.from_range_0_1_to_parameter_value(np.random.beta(1, 1, size))

    randomly_select = randomly_select_uniform


class IntegerParam(_Param):  # space.py:345-465
    kind = "integer"

    def __init__(self, lo, hi, default=None):
        if lo > hi:
            lo, hi = hi, lo
        self.min_value, self.max_value, self.default = lo, hi, default
        self.values_list = list(range(lo, hi + 1))

    def get_size(self):
        return self.max_value - self.min_value + 1

    get_discrete_size = get_size

    def get_discrete_values(self):
        return range(self.min_value, self.max_value + 1)

    def get_min(self):
        return self.min_value

    def get_max(self):
        return self.max_value

    def from_range_0_1_to_parameter_value(self, X):   # space.py:437-453
        Xc = np.array([X]).copy() if type(X) != np.ndarray and type(X) != list else np.array(X).copy()
        s = np.round(Xc * self.get_size() - 0.5 + self.min_value).astype(int)
        return s[0] if len(s) == 1 else s

    def randomly_select_uniform(self, size=1):        # space.py:389-395
        return np.random.choice(self.values_list, size=size).astype(int)

    randomly_select = randomly_select_uniform


class OrdinalParam(_Param):  # space.py:467-596
    kind = "ordinal"

    def __init__(self, values, default=None):
        self.values = sorted(values, key=float)
        self.default = default

    def get_size(self):
        return len(self.values)

    get_discrete_size = get_size

    def get_discrete_values(self):
        return self.values

    get_values = get_discrete_values

    def get_min(self):
        return self.values[0]

    def get_max(self):
        return self.values[-1]

    def from_range_0_1_to_parameter_value(self, X):   # space.py:578-595
        Xc = np.array([X]).copy() if type(X) != np.ndarray and type(X) != list else np.array(X).copy()
        s = np.round(Xc * self.get_size() - 0.5).astype(int)
        return self.values[s[0]] if len(s) == 1 else np.array(self.values)[s]

    def randomly_select_uniform(self, size=1):        # space.py:505-511
        return np.random.choice(self.values, size=size)

    randomly_select = randomly_select_uniform


class CategoricalParam(_Param):  # space.py:598-697 -- values are handled as integer indices everywhere
    kind = "categorical"

    def __init__(self, values, default=None):
        self.values = values
        self.default = default

    def get_size(self):
        return len(self.values)

    get_discrete_size = get_size

    def get_discrete_values(self):
        return list(range(len(self.values)))

    def get_values(self):
        return self.values

    def get_default(self):
        return None if self.default is None else self.values.index(self.default)

    def get_int_value(self, s):
        return self.values.index(s)

    def get_original_string(self, i):
        return self.values[i]

    def randomly_select_uniform(self, size=1):        # space.py:639-645
        return np.random.choice(self.get_size(), size=size)

    randomly_select = randomly_select_uniform


class SpaceLite:
    def __init__(self, config):
        self.optimization_metrics = list(config["optimization_objectives"])
        fo = config.get("feasible_output", {})
        self.enable_feasible_predictor = bool(fo.get("enable_feasible_predictor", False))
        self.feasible_output_name = fo.get("name", "Valid") if self.enable_feasible_predictor else None
        self.normalize_inputs = bool(config.get("normalize_inputs", False))
        self.parameter_means, self.parameter_stds = {}, {}
        self.all_input_parameters = OrderedDict()
        self.parameters_type = OrderedDict()
        for name, p in config["input_parameters"].items():
            t = p["parameter_type"]
            d = p.get("parameter_default")
            if t == "real":
                obj = RealParam(p["values"][0], p["values"][1], d)
            elif t == "integer":
                obj = IntegerParam(p["values"][0], p["values"][1], d)
            elif t == "ordinal":
                obj = OrdinalParam(p["values"], d)
            elif t == "categorical":
                obj = CategoricalParam(p["values"], d)
            else:
                print("Unsupported parameter type:", t)
                raise SystemExit
            self.all_input_parameters[name] = obj
            self.parameters_type[name] = t

    # --- accessors (space.py:1022-1186) ---
    def get_type(self, parameter):
        return self.parameters_type[parameter]

    def get_input_parameters(self):
        return list(self.all_input_parameters.keys())

    def get_input_parameters_objects(self):
        return self.all_input_parameters

    def get_input_categorical_parameters(self, names):
        return [p for p in names if self.parameters_type.get(p) == "categorical"]

    def get_input_non_categorical_parameters(self, names):
        return [p for p in names if p in self.parameters_type and self.parameters_type[p] != "categorical"]

    def get_optimization_parameters(self):
        return self.optimization_metrics

    def get_feasible_parameter(self):
        return [self.feasible_output_name]

    def get_input_normalization_flag(self):
        return self.normalize_inputs

    def get_parameter_mean(self, p):
        return self.parameter_means[p]

    def get_parameter_std(self, p):
        return self.parameter_stds[p]

    def set_parameter_mean(self, p, v):
        self.parameter_means[p] = v

    def set_parameter_std(self, p, v):
        self.parameter_stds[p] = v

    def get_space_size(self):              # space.py:1186-1194
        total = 1
        for p in self.all_input_parameters.values():
            total *= p.get_size()
        return total

    def get_space(self):                   # space.py:1222-1241
        prod = itertools.product(*[p.get_discrete_values() for p in self.all_input_parameters.values()])
        out = defaultdict(list)
        names = self.get_input_parameters()
        for conf in prod:
            for j, h in enumerate(names):
                out[h].append(conf[j])
        return out

    def get_random_configuration(self, use_priors=True, size=1, return_as_array=False):   # space.py:1263-1326
        names = self.get_input_parameters()
        samples = np.array([[None] * len(names)] * size)
        for i, k in enumerate(names):
            samples[:, i] = self.all_input_parameters[k].randomly_select_uniform(size=size)
        if return_as_array:
            return samples
        if size == 1:
            return {k: samples[0][i] for i, k in enumerate(names)}
        return [{k: samples[r, i] for i, k in enumerate(names)} for r in range(size)]

    def get_default_configuration(self):
        return {k: p.get_default() for k, p in self.all_input_parameters.items()}

    def get_default_or_random_configuration(self, user_priors=False):   # space.py:1347-1359
        conf = self.get_default_configuration()
        if None in conf.values():
            rnd = self.get_random_configuration(use_priors=user_priors)
            for k in conf:
                if conf[k] is None:
                    conf[k] = rnd[k]
        return conf

    def get_unique_hash_string_from_values(self, configuration):        # space.py:1699-1720
        s = ""
        for p in self.get_input_parameters():
            s += (str(configuration[p]) if p in configuration else "") + "_"
        return s

    def remove_duplicate_configs(self, *configs, ignore_columns=None):  # space.py:1648-1678
        full = np.concatenate(configs, axis=0).astype(float)
        if ignore_columns is not None:
            full = np.delete(full, ignore_columns, axis=1)
        _, uidx = np.unique(full, return_index=True, axis=0)
        starts = np.cumsum(np.insert(np.array([len(c) for c in configs]), 0, 0))
        out = []
        for i, c in enumerate(configs):
            sel = uidx[(uidx >= starts[i]) & (uidx < starts[i + 1])] - starts[i]
            out.append(c[sel])
        return out
