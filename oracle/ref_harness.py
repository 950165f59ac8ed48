"""Import the UNMODIFIED reference (luinardi/hypermapper under /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py to generate golden vectors from the
real reference; it cannot travel to the GPU box (/root/reference does not exist there), so nothing in
the `-m gpu` tests, smoke() or bench.py imports this module.

The reference imports GPy / cma / matplotlib eagerly (hypermapper/__init__.py:4-18); they are not
installed here, so they are replaced by empty stub modules.  `"-m"` in sys.argv makes
hypermapper/__init__.py:4 skip the eager imports.  models.py:977 passes OneHotEncoder(sparse=False),
removed in scikit-learn >= 1.4, so a one-line shim maps it to sparse_output.  Hot-path code calls
sys.stdout.write_to_logfile (models.py:345,493; local_search.py:454,...), so a Logger-like stdout
wrapper is installed while reference code runs.
"""
import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HM_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hypermapper"))


class _LoggerStdout:
    """Duck-typed stand-in for utility_functions.Logger (utility_functions.py:116-181): a stdout
    that also accepts write_to_logfile / write_protocol / switch_log_only_on_file."""

    def __init__(self, inner):
        self._inner = inner

    def write(self, msg):
        return self._inner.write(msg)

    def flush(self):
        return self._inner.flush()

    def write_to_logfile(self, msg, msg_is_verbose=False):
        pass

    def write_protocol(self, msg):
        pass

    def switch_log_only_on_file(self, choice):
        pass

    def __getattr__(self, name):
        return getattr(self._inner, name)


@contextlib.contextmanager
def logger_stdout():
    old = sys.stdout
    sys.stdout = _LoggerStdout(old)
    try:
        yield
    finally:
        sys.stdout = old


_loaded = None


def load():
    """Returns a namespace with the reference modules: models, random_scalarizations, local_search,
    compute_pareto, utility_functions, space, prior_optimization, pibo."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("GPy", "cma", "matplotlib", "matplotlib.pyplot", "matplotlib.lines", "matplotlib.cm",
                 "pygmo", "statsmodels", "statsmodels.api", "pyDOE"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "-m" not in sys.argv:
        sys.argv.append("-m")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    ns = types.SimpleNamespace()
    for mod in ("utility_functions", "space", "models", "local_search", "random_scalarizations", "compute_pareto",
                "prior_optimization", "pibo"):
        setattr(ns, mod, importlib.import_module("hypermapper." + mod))
    import sklearn.preprocessing as skp

    def _ohe(categories="auto", sparse=False, **kw):
        return skp.OneHotEncoder(categories=categories, sparse_output=sparse, **kw)

    ns.models.OneHotEncoder = _ohe
    _loaded = ns
    return ns
