"""Scaffolding for bench.py, smoke() and the tests -- NOT part of the product package.

The hot path takes a fitted surrogate and a search-space object from the reference (model fitting and `space.Space` are
out of scope).  Neither travels to the GPU box, so this package provides stand-ins:
    space_lite.SpaceLite      the accessor surface of space.Space the hot path duck-types against
    hutter_forest             sklearn fit + Hutter per-leaf tables + re-drawn thresholds, as RFModel.fit_RF does
    workloads                 the scenarios of BASELINE.json (C1, C3, C5, Hartmann-6) with seeded data
"""
