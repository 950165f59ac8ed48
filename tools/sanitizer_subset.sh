#!/bin/bash
# compute-sanitizer record: memcheck / racecheck / synccheck over a small-shape subset of the GPU
# tests that reaches every hand-rolled synchronisation pattern of the library:
#   forest kernel   streamed through the TMA + mbarrier double buffer, resident, walked in L2; coherence pre-pass
#   top-k           shared-atomic candidate buffer with roll-back (k = 20 and k = 1024)
#   GP kernel       4- and 8-warp plans, per-warp TMA rings
#   Pareto          fast path (warp-autonomous queues) and exact NaN path
#   hill climb      hm_ls_neighbors / hm_ls_move
# usage (on the GPU box): tools/sanitizer_subset.sh <outdir>
OUT=${1:-gpurun_out}
IDS=(
 "tests/test_rf_gpu.py::test_forest_kernel_vs_oracle_synthetic"
 "tests/test_rf_gpu.py::test_coherence_prepass_is_invisible[24-32-200001]"
 "tests/test_rf_gpu.py::test_topk_vs_oracle_ties_nan[100000-20]"
 "tests/test_rf_gpu.py::test_topk_vs_oracle_ties_nan[3000000-1024]"
 "tests/test_rf_gpu.py::test_topk_descending_input_worst_case"
 "tests/test_rf_gpu.py::test_forest_edge_cases_empty_single_leaf_mixed_layouts_eight_objectives"
 "tests/test_gp_gpu.py::test_gp_mean_var_vs_sklearn_golden"
 "tests/test_gp_gpu.py::test_gp_mean_var_vs_oracle[17-3-0.01-1000-matern52]"
 "tests/test_pareto_gpu.py::test_mask_bit_exact_vs_reference_golden"
 "tests/test_pareto_gpu.py::test_vs_oracle_seeded[50000-3-u]"
 "tests/test_pareto_gpu.py::test_vs_oracle_seeded[40000-3-nan]"
 "tests/test_pareto_gpu.py::test_vs_oracle_seeded[200000-3-dups]"
 "tests/test_space_gpu.py::test_device_neighbourhoods_equal_get_neighbors"
 "tests/test_space_gpu.py::test_device_encode_matches_reference_preprocess"
 "tests/test_api_gpu.py::test_lockstep_hill_climb_equals_sequential"
)
rc_all=0
for tool in memcheck racecheck synccheck; do
  t0=$(date +%s)
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --log-file $OUT/r02_sanitizer_$tool.log \
      python -m pytest "${IDS[@]}" -x -q -p no:cacheprovider > $OUT/r02_sanitizer_${tool}_pytest.log 2>&1
  rc=$?
  echo "$tool rc=$rc seconds=$(( $(date +%s) - t0 )) $(tail -1 $OUT/r02_sanitizer_${tool}_pytest.log)" | tee -a $OUT/r02_sanitizer_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/r02_sanitizer_$tool.log | tail -2 | tee -a $OUT/r02_sanitizer_summary.txt
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
