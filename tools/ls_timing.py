"""Wall-clock of one acquisition-optimisation iteration through the public API (random_scalarizations -> local_search)
on the C1 (Branin, 10-tree RF) and C3-like scenarios at the reference's default pool sizes.

    python tools/ls_timing.py [--points 10000] [--starts 10] [--reference]

--reference times the UNMODIFIED reference on the same fitted models (build container only: needs /root/reference)."""
import argparse
import copy
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=10000)
    ap.add_argument("--starts", type=int, default=10)
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--scenario", default="c1", choices=["c1", "c3"])
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    from bench_support import workloads
    scen = workloads.scenario_c1() if args.scenario == "c1" else workloads.scenario_c3()
    if args.scenario == "c3":
        scen["models"]["number_of_trees"] = 32
    wl = workloads.RFWorkload(scen, n_train=15 if args.scenario == "c1" else 300, seed=0)
    space = wl.space
    inputs = space.get_input_parameters()
    data = {p: [v.item() if hasattr(v, "item") else v for v in wl.train_cols[p]] for p in inputs}
    for o in wl.objectives:
        data[o] = list(wl.data_array[o])
    config = {"models": {"model": "random_forest"}, "acquisition_function": "EI", "scalarization_method": "tchebyshev",
              "number_of_cpus": 1, "local_search_starting_points": args.starts,
              "local_search_random_points": args.points, "scalarization_key": "scalarization",
              "optimization_objectives": list(wl.objectives)}
    fast = {space.get_unique_hash_string_from_values({p: data[p][r] for p in inputs}): r
            for r in range(len(data[inputs[0]]))}

    def run(rs_mod, uf_mod, models, sp):
        d2 = copy.deepcopy(data)
        scal, _ = uf_mod.compute_data_array_scalarization(d2, wl.objective_weights, copy.deepcopy(wl.objective_limits),
                                                          "tchebyshev")
        d2["scalarization"] = list(scal)
        np.random.seed(1)
        random.seed(1)
        t0 = time.perf_counter()
        best = rs_mod.random_scalarizations(config, d2, sp, fast, models, 5, wl.objective_weights,
                                            copy.deepcopy(wl.objective_limits), None, None, "local_search")
        return time.perf_counter() - t0, best

    if args.reference:
        from oracle import ref_harness
        ns = ref_harness.load()
        import json
        cfg = dict(scen)
        cfg.setdefault("application_name", "x")
        from jsonschema import Draft4Validator
        schema = json.load(open(os.path.join(ref_harness.REFERENCE_ROOT, "hypermapper", "schema.json")))
        V = ns.utility_functions.extend_with_default(Draft4Validator)
        V(schema).validate(cfg)
        rspace = ns.space.Space(cfg)
        # the reference needs RFModel objects: wrap the fitted forests
        rmodels = {}
        for o, m in wl.models.items():
            rf = ns.models.RFModel(n_estimators=len(m.estimators_))
            rf.__dict__.update({k: v for k, v in m.__dict__.items()})
            rmodels[o] = rf
        with ref_harness.logger_stdout():
            for _ in range(args.reps):
                dt, best = run(ns.random_scalarizations, ns.utility_functions, rmodels, rspace)
                print("reference  %-3s N=%d starts=%d: %.3f s  best=%s" % (args.scenario, args.points, args.starts, dt, best))
        return
    from hypermapper_b200 import random_scalarizations as rs, utility_functions as uf
    for _ in range(args.reps):
        dt, best = run(rs, uf, wl.models, space)
        print("b200       %-3s N=%d starts=%d: %.3f s  best=%s" % (args.scenario, args.points, args.starts, dt, best))


if __name__ == "__main__":
    main()
