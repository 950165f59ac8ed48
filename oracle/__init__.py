"""CPU oracle: a restatement of the reference's algorithm for the hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product (hypermapper_b200/) never does and fails
loudly when its CUDA library is missing.

Pinning status
  * RF leaves / mean / variance / sklearn predict, EI / UCB / TS x {linear, tchebyshev,
    modified_tchebyshev}, encoding, get_min_configurations, compute_data_array_scalarization,
    local_search seed selection, Pareto filter: PINNED -- tests/test_oracle.py checks them against
    tests/golden/*.npz, which tests/golden/make_golden.py produced by running the UNMODIFIED reference
    (/root/reference, luinardi/hypermapper v2.2.12) in the build container, and against the three
    Pareto known-answer CSVs of the reference's own tests (tests/data/*pareto*.csv, re-derived).
  * GP posterior mean/variance: PARITY UNPINNED.  The reference delegates to the third-party
    package GPy (setup.py:29 unpinned, .travis.yml:23 pins GPy==1.9.9) which is not vendored under
    /root/reference and not installable here (no network).  oracle.gp_predict restates GPy 1.9.9's
    published exact-inference algorithm (Matern52 ARD kernel, Standardize normaliser, pdinv/dpotrs
    fit, dtrtrs predict) and is anchored on the reference's call sites (models.py:444-457, 1011-1018).
"""
from .oracle import *  # noqa: F401,F403
