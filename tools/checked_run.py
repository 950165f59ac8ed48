#!/usr/bin/env python
"""Runs a small-shape sweep of every kernel family against the CHECKED build of the library (make -C
hypermapper_b200/csrc checked) and reports hm_checked_failures(): the substitute for a compute-sanitizer record on a
GPU pool where compute-sanitizer is closed (profiles/r02_sanitizer_unavailable.txt).

    HM_B200_LIB=hypermapper_b200/csrc/libhm_b200_checked.so python tools/checked_run.py

Coverage: forest kernel (resident with leaf records in shared memory / streamed through the TMA double buffer + coherence pre-pass / walked in L2; fp32,
packed-fast, packed-wide and self-looping-leaf layouts, malformed codes), top-k (k = 20 and k = 1024, heavy ties: the roll-back path), GP (4- and 8-warp
plans, per-warp TMA rings; the large-n path), Pareto (fast path with warp queues, exact NaN path), encode / sample,
device hill climb.  Every case also checks its results against the oracle / goldens (they are the GPU tests' own
functions), so a wrong result fails here too."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from hypermapper_b200 import _lib
    L = _lib.lib()
    tu, line = ctypes.c_int(-1), ctypes.c_int(0)
    assert L.hm_checked_failures(ctypes.byref(tu), ctypes.byref(line)) == 0, \
        "not a checked build (HM_B200_LIB=%s)" % os.environ.get("HM_B200_LIB")
    import test_rf_gpu as rf
    import test_gp_gpu as gp
    import test_pareto_gpu as pa
    import test_packed_gpu as pk
    import test_space_gpu as sp
    import test_api_gpu as api
    import test_selection_gpu as sel
    cases = [
        (rf.test_forest_kernel_vs_oracle_synthetic, (3, 2, 1, 3, 5000)),
        (rf.test_forest_kernel_vs_oracle_synthetic, (64, 23, 4, 11, 20000)),
        (rf.test_forest_kernel_vs_oracle_synthetic, (300, 32, 6, 12, 6000)),
        (rf.test_forest_kernel_vs_oracle_synthetic, (4, 5, 13, 15, 4000)),
        (rf.test_coherence_prepass_is_invisible, (24, 32, 200_001)),
        (rf.test_forest_edge_cases_empty_single_leaf_mixed_layouts_eight_objectives, ()),
        (rf.test_resident_forests_with_leaf_records_in_shared_memory, ()),
        (pk.test_packed_path_is_bit_identical_to_fp32_path, ("c3", 5000, 7, 3, 9)),
        (rf.test_topk_vs_oracle_ties_nan, (100000, 20)),
        (rf.test_topk_vs_oracle_ties_nan, (3_000_000, 1024)),
        (rf.test_topk_descending_input_worst_case, ()),
        (rf.test_topk_merge_equals_global_topk, ()),
        (pk.test_packed_path_is_bit_identical_to_fp32_path, ("c3", 300_000, 40, 8, 12)),
        (pk.test_packed_path_is_bit_identical_to_fp32_path, ("odd", 4000, 5, 2, 8)),
        (pk.test_packed_path_is_bit_identical_to_fp32_path, ("wide", 150_000, 40, 8, 11)),
        (pk.test_packed_path_is_bit_identical_to_fp32_path, ("c3", 3000, 3, 13, 15)),
        (pk.test_packed_path_is_bit_identical_to_fp32_path, ("pow2", 140_000, 40, 8, 11)),
        (pk.test_malformed_codes_stay_inside_the_forest, ("c3", 200_000, 24)),
        (pk.test_malformed_codes_stay_inside_the_forest, ("odd", 50_000, 12)),
        (pk.test_malformed_codes_stay_inside_the_forest, ("pow2", 50_000, 12)),
        (pk.test_malformed_codes_stay_inside_the_forest, ("wide", 50_000, 12)),
        (pk.test_packed_host_entry_and_errors, ()),
        (sel.test_fused_score_topk_equals_eval_then_topk, ("c3_packed",)),
        (sel.test_fused_score_topk_equals_eval_then_topk, ("overflow",)),
        (gp.test_gp_mean_var_vs_oracle, (256, 6, 1e-4, 20000, "matern52")),
        (gp.test_gp_mean_var_vs_oracle, (512, 32, 1e-4, 4000, "matern52")),
        (gp.test_gp_mean_var_vs_oracle, (700, 10, 1e-3, 2000, "matern52")),
        (gp.test_gp_mean_var_vs_oracle, (17, 3, 1e-2, 1000, "matern52")),
        (gp.test_gp_large_models_beyond_the_fused_kernel, (1000, 8, 1e-3, 3000)),
        (pa.test_vs_oracle_seeded, (50000, 3, "u")),
        (pa.test_vs_oracle_seeded, (40000, 3, "nan")),
        (pa.test_vs_oracle_seeded, (200000, 3, "dups")),
        (pa.test_vs_oracle_seeded, (20000, 8, "u")),
        (pa.test_full_size_properties_10M_x3, ()),
        (sp.test_sample_pool_distributions_and_regeneration, ()),
        (sp.test_device_neighbourhoods_equal_get_neighbors, ()),
        (api.test_lockstep_hill_climb_equals_sequential, ()),
    ] + [(pa.test_mask_bit_exact_vs_reference_golden, (c,)) for c in pa.CASES]
    for fn, args in cases:
        fn(*args)
        n = L.hm_checked_failures(ctypes.byref(tu), ctypes.byref(line))
        print("%-70s %-40s violations so far: %d" % (fn.__name__, str(args)[:40], n), flush=True)
    n = L.hm_checked_failures(ctypes.byref(tu), ctypes.byref(line))
    names = ["hm_forest.cu", "hm_acq.cu", "hm_gp.cu", "hm_pareto.cu", "hm_space.cu"]
    if n:
        print("CHECKED BUILD: %d invariant violations, last in %s line %d" % (n, names[tu.value], line.value))
        sys.exit(1)
    print("CHECKED BUILD: %d cases, 0 invariant violations" % len(cases))


if __name__ == "__main__":
    main()
