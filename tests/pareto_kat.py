"""The two grid known-answer cases of the reference's own tests (tests/test_system.py:36-64): the exhaustive grids are
regenerated here (the objective is example_scenarios/synthetic/multiobjective_branin/multiobjective_branin.py:13-36,
restated with the same scalar `math` arithmetic), the expected fronts are the rows of the reference's
tests/data/branin_grid_search_pareto.csv (98 rows) and tests/data/ordinal_branin_real_pareto.csv (12 rows), frozen in
tests/golden/pareto.npz by make_golden.py."""
import json
import math

import numpy as np


def branin_two_objectives(x1, x2):
    a = 1.0
    b = 5.1 / (4.0 * math.pi * math.pi)
    c = 5.0 / math.pi
    r = 6.0
    s = 10.0
    t = 1.0 / (8.0 * math.pi)
    y_value = a * (x2 - b * x1 * x1 + c * x1 - r) ** 2 + s * (1 - t) * math.cos(x1) + s
    return y_value, x1 + x2


def grid(kind):
    """-> rows [n][4] = x1, x2, Value, Energy of the exhaustive grid"""
    if kind == "branin_grid":
        xs1, xs2 = np.linspace(-5, 10, 100).tolist(), np.linspace(0, 15, 100).tolist()
    else:
        xs1, xs2 = [float(v) for v in range(-5, 11)], [float(v) for v in range(0, 16)]
    rows = []
    for x1 in xs1:
        for x2 in xs2:
            v, e = branin_two_objectives(x1, x2)
            rows.append([x1, x2, v, e])
    return np.array(rows, dtype=np.float64)


def check(z, kind, mask_fn):
    """mask_fn(costs [n][2]) -> bool mask; the front must be exactly the reference's CSV rows"""
    rows = grid(kind)
    mask = np.asarray(mask_fn(np.ascontiguousarray(rows[:, 2:4]))).astype(bool)
    cols = json.loads(str(z["kat_%s_cols" % kind]))
    kat = z["kat_" + kind][:, [cols.index(c) for c in ("x1", "x2", "Value", "Energy")]]
    got = np.array(sorted(rows[mask].tolist()))
    want = np.array(sorted(kat.tolist()))
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got[:, :2], want[:, :2])                       # the same grid points ...
    assert np.allclose(got[:, 2:], want[:, 2:], rtol=1e-12, atol=0)      # ... with the CSV's objective values
