"""A small stand-in for the reference's `space.Space` exposing only the accessor surface the hot path uses.

The search-space model (space.py, 2774 lines: priors, KDEs, CSV IO, black-box dispatch ...) is OUT OF SCOPE and
stays with the reference; in a deployment the real `Space` object is passed in and every function of this
package duck-types against it.  `SpaceLite` exists so that tests, smoke() and bench.py can run on a box where
the reference is not installed.  It is built from the same scenario JSON ("input_parameters",
"optimization_objectives", "feasible_output", "normalize_inputs"; schema.json is unchanged) and reproduces the
accessor semantics cited below, including the per-parameter priors ("prior": a named Beta density, a probability
list, or "custom_gaussian"; the multivariate KDE "estimate" prior is not modelled).
"""
import itertools
from collections import OrderedDict, defaultdict

import numpy as np
from scipy import stats

# space.py:74-85
DENSITIES_ALPHAS = {"uniform": 1.0, "gaussian": 3.0, "decay": 0.8, "exponential": 1}
DENSITIES_BETAS = {"uniform": 1.0, "gaussian": 3.0, "decay": 1, "exponential": 0.8}


class _Param:
    kind = None
    densities_alphas = DENSITIES_ALPHAS
    densities_betas = DENSITIES_BETAS

    def get_default(self):
        return self.default

    def get_prior(self):
        return self.prior


class RealParam(_Param):  # space.py:88-343
    kind = "real"

    def __init__(self, lo, hi, default=None, prior="uniform"):       # space.py:93-130
        self.min_value, self.max_value, self.default = lo, hi, default
        self.preferred_discretization = np.linspace(lo, hi, num=10)
        self.prior = prior
        self.alpha = self.beta = None
        self.custom_gaussian_prior_mean = self.custom_gaussian_prior_std = None
        if isinstance(prior, str) and prior in self.densities_alphas:
            self.alpha, self.beta = self.densities_alphas[prior], self.densities_betas[prior]
        self.cdf_distribution = None
        if isinstance(prior, list):
            self.prior = list(prior / np.sum(prior))
            self.param_xs = np.linspace(lo, hi, 10000)
            self.pdf_distribution = self.interpolate_pdf(self.param_xs)
            self.cdf_distribution = np.cumsum(self.pdf_distribution)
            self.cdf_distribution = self.cdf_distribution / self.cdf_distribution[-1]

    def set_custom_gaussian_prior_params(self, mean, std):           # space.py:306-314
        if std == -1:
            std = (self.max_value - self.min_value) / 2
        self.custom_gaussian_prior_mean, self.custom_gaussian_prior_std = mean, std

    def interpolate_pdf(self, x):                                    # space.py:222-229
        return np.interp(x, np.linspace(self.min_value, self.max_value, len(self.prior)), self.prior)

    def from_parameter_value_to_0_1_range(self, X):                  # space.py:333-342
        return (X - self.min_value) / (self.max_value - self.min_value)

    def get_x_probability(self, x):                                  # space.py:264-286
        if isinstance(self.prior, list):
            return self.interpolate_pdf(x)
        if self.prior == "custom_gaussian":
            return stats.norm.pdf(x, loc=self.custom_gaussian_prior_mean, scale=self.custom_gaussian_prior_std)
        pdf = stats.beta.pdf(self.from_parameter_value_to_0_1_range(x), self.alpha, self.beta)
        return 1242 if pdf == float("inf") else pdf

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

    def randomly_select(self, size=1):                # space.py:138-199
        if isinstance(self.prior, list):
            import random
            return np.interp(random.uniform(0, 1), self.cdf_distribution, self.param_xs)
        if self.prior == "custom_gaussian":
            samples = np.array([])
            while len(samples) < size:
                new = stats.norm.rvs(loc=self.custom_gaussian_prior_mean, scale=self.custom_gaussian_prior_std,
                                     size=size * 5)
                samples = np.append(samples, new[(new <= self.max_value) & (new >= self.min_value)])
            return samples[0:size]
        return self.from_range_0_1_to_parameter_value(np.random.beta(self.alpha, self.beta, size))


class IntegerParam(_Param):  # space.py:345-465
    kind = "integer"

    def __init__(self, lo, hi, default=None, prior="uniform"):       # space.py:350-370
        if lo > hi:
            lo, hi = hi, lo
        self.min_value, self.max_value, self.default = lo, hi, default
        self.values_list = list(range(lo, hi + 1))
        self.distribution = None
        if isinstance(prior, str):
            self.prior = prior
            self.alpha, self.beta = self.densities_alphas[prior], self.densities_betas[prior]
        else:
            self.prior, self.distribution = "distribution", prior

    def get_x_probability(self, x):                                  # space.py:397-417
        if self.prior == "distribution":
            return self.distribution[self.values_list.index(x)]
        return stats.beta.pdf((x - self.min_value) / (self.max_value - self.min_value), self.alpha, self.beta)

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

    def randomly_select(self, size=1):                # space.py:372-387
        if self.prior == "distribution":
            return np.random.choice(self.values_list, size=size, p=self.distribution).astype(int)
        return self.from_range_0_1_to_parameter_value(np.random.beta(self.alpha, self.beta, size))


class OrdinalParam(_Param):  # space.py:467-596
    kind = "ordinal"

    def __init__(self, values, default=None, prior="uniform"):       # space.py:472-487
        self.values = sorted(values, key=float)
        self.default = default
        self.distribution = []
        if isinstance(prior, str):
            self.prior = prior
        else:
            self.prior, self.distribution = "distribution", prior

    def get_parameter_distribution(self):
        return self.distribution

    def get_x_probability(self, x):                                  # space.py:513-540
        i = self.values.index(x)
        if self.prior == "distribution":
            return self.distribution[i]
        a, b = self.densities_alphas[self.prior], self.densities_betas[self.prior]
        return stats.beta.cdf((i + 1) / len(self.values), a, b) - stats.beta.cdf(i / len(self.values), a, b)

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

    def randomly_select(self, size=1):                # space.py:489-503
        if self.prior == "distribution":
            return np.random.choice(self.values, size=size, p=self.distribution)
        return self.from_range_0_1_to_parameter_value(
            np.random.beta(self.densities_alphas[self.prior], self.densities_betas[self.prior], size))


class CategoricalParam(_Param):  # space.py:598-697 -- values are handled as integer indices everywhere
    kind = "categorical"

    def __init__(self, values, default=None, prior="uniform"):       # space.py:606-625
        self.values = values
        self.default = default
        if isinstance(prior, str):
            if prior != "uniform":
                print("Error in the json file: only uniform or a list of probabilities is valid for categorical "
                      "parameters. Exit.")
                raise SystemExit
            self.prior = "uniform"
            self.distribution = [1.0 / len(values)] * len(values)
        else:
            self.prior, self.distribution = "distribution", prior

    def get_parameter_distribution(self):
        return self.distribution

    def get_x_probability(self, x):                                  # space.py:647-660
        return self.distribution[x]

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

    def randomly_select(self, size=1):                # space.py:627-637
        if self.prior == "uniform":
            return np.random.choice(self.get_size(), size=size)
        return np.random.choice(self.get_size(), size=size, p=self.distribution)


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
            prior = p.get("prior", "uniform")
            if t == "real":
                obj = RealParam(p["values"][0], p["values"][1], d, prior)
                if prior == "custom_gaussian":     # space.py:878-907 (one mean/std, or one per input parameter)
                    means, stds = config["custom_gaussian_prior_means"], config["custom_gaussian_prior_stds"]
                    j = list(config["input_parameters"]).index(name)
                    obj.set_custom_gaussian_prior_params(means[0] if len(means) == 1 else means[j],
                                                         stds[0] if len(stds) == 1 else stds[j])
            elif t == "integer":
                obj = IntegerParam(p["values"][0], p["values"][1], d, prior)
            elif t == "ordinal":
                obj = OrdinalParam(p["values"], d, prior)
            elif t == "categorical":
                obj = CategoricalParam(p["values"], d, prior)
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

    def get_estimate_prior_flags(self):    # space.py:1173-1178
        return [False]

    def get_prior_normalization_flag(self):   # space.py:763, 830-911, 1407-1411: any named (string) prior sets it
        return any(isinstance(getattr(p, "prior", None), str) and not isinstance(p, CategoricalParam) or
                   (isinstance(p, CategoricalParam) and p.prior == "uniform")
                   for p in self.all_input_parameters.values())

    def random_sample_configurations_without_repetitions(self, fast_addressing_of_data_array, number_of_RS,
                                                         use_priors=True):   # space.py:1797-1860 (large-space branch)
        configurations = []
        arr = list(fast_addressing_of_data_array.values())
        index = 0 if len(arr) == 0 else int(np.asarray(arr, dtype=np.int64).max()) + 1
        tries = 0
        while len(configurations) != number_of_RS and tries < number_of_RS + 1000000:
            tries += 1
            configuration = self.get_random_configuration(use_priors)
            key = self.get_unique_hash_string_from_values(configuration)
            if key in fast_addressing_of_data_array:
                continue
            fast_addressing_of_data_array[key] = index
            index += 1
            configurations.append(configuration)
        return configurations

    def get_discrete_space_size(self):     # space.py:1196-1206
        total = 1
        for p in self.all_input_parameters.values():
            total *= p.get_discrete_size()
        return total

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
            if use_priors:
                samples[:, i] = self.all_input_parameters[k].randomly_select(size=size)
            else:
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
