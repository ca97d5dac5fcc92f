#!/bin/bash
# compute-sanitizer over the tcgen05 kernels (small shapes): memcheck on every fast path, then racecheck and synccheck on one test
# per kernel (hand-rolled mbarrier protocols).  Summaries -> gpurun_out/sanitize_summary.txt (copied to profiles/r02/).
# Every error line is counted and the first ones are printed, whatever their kind (round 1 grepped only for three kinds).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
out=gpurun_out/sanitize_summary.txt; : > $out
run() {   # tool, tag, pytest -k expression
  local tool=$1 tag=$2 expr=$3 log=gpurun_out/san_${1}_${2}.log
  timeout ${SAN_TIMEOUT:-400} compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "$expr" > $log 2>&1
  local rc=$?
  {
    echo "== $tool / $tag  (pytest -k \"$expr\")  rc=$rc"
    echo "   $(grep -a -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
    echo "   pytest: $(grep -a -E '[0-9]+ (passed|failed)' $log | tail -1)"
    grep -a -E '^=========     (Invalid|Out-of-range|Misaligned|Race|Barrier|Uninitialized|Program hit|Error:|Warning:|Potential|Divergent|Hazard)' $log | sort | uniq -c | sort -rn | head -6 | sed 's/^/   /'
  } >> $out
}
run memcheck fused28   "test_fused_kernel_many_items_per_cta_fp16 and 28-4-120"
run memcheck cluster   "test_fused_kernel_many_items_per_cta_fp16 and 28-4-200"
run memcheck fused14   "test_fused_kernel_many_items_per_cta_fp16 and 14-2"
run memcheck causal    "test_causal_tcgen05_many_windows_per_cta_fp16"
run memcheck lara      "test_lara_core_many_items_vs_oracle_fp16"
run memcheck window_tc "test_tcgen05_window_kernel_matches_the_cuda_core_kernel"
run memcheck bwd_tc    "test_tcgen05_backward_kernel_matches_the_cuda_core_kernel"
run memcheck bwd_simt  "test_backward_kernels_equal_autograd_through_the_recomputation"
run memcheck bwd_gen   "test_generic_tcgen05_backward_kernel_matches_the_cuda_core_kernel"
run memcheck stats_fast "test_fast_chunk_statistics_kernel_matches_the_generic_one"
run memcheck lara_bwd  "test_lara_fused_backward_steps_match_the_explicit_formulas"
run racecheck fused28  "test_fused_kernel_variants_fp16 and 28-4-False and default and True-True"
run racecheck cluster  "test_fused_kernel_variants_fp16 and 28-4-True and default and True-True"
run racecheck causal   "test_causal_tcgen05_window_kernel_vs_oracle and 256-True-True and dtype0"
run racecheck lara     "test_lara_core_only_prequantised_fp16_vs_oracle and True-False"
run racecheck window_tc "test_tcgen05_window_kernel_matches_the_cuda_core_kernel and seq_shape3"
run synccheck window_tc "test_tcgen05_window_kernel_matches_the_cuda_core_kernel and seq_shape3"
run racecheck bwd_tc   "test_tcgen05_backward_kernel_matches_the_cuda_core_kernel and seq_shape0"
run racecheck bwd_simt "test_backward_kernels_equal_autograd_through_the_recomputation and seq_shape2"
run racecheck bwd_gen  "test_generic_tcgen05_backward_kernel_matches_the_cuda_core_kernel and seq_shape2"
run racecheck lara_bwd "test_lara_fused_backward_steps_match_the_explicit_formulas and dtype1"
run synccheck bwd_gen  "test_generic_tcgen05_backward_kernel_matches_the_cuda_core_kernel and seq_shape2"
run synccheck bwd_tc   "test_tcgen05_backward_kernel_matches_the_cuda_core_kernel and seq_shape0"
run synccheck fused28  "test_fused_kernel_variants_fp16 and 28-4-False and default and True-True"
run synccheck cluster  "test_fused_kernel_variants_fp16 and 28-4-True and default and True-True"
run synccheck causal   "test_causal_tcgen05_window_kernel_vs_oracle and 256-True-True and dtype0"
run synccheck lara     "test_lara_core_only_prequantised_fp16_vs_oracle and True-False"
cat $out
