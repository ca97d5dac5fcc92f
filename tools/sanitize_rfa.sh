#!/bin/bash
# compute-sanitizer over the random-feature kernels of the round (rfa_kernels.cu, rfa_tc_sm100.cu, sb_window_tc_sm100.cu and the
# key-bias path of eva_window_tc_sm100.cu): memcheck, racecheck, synccheck.  Summary -> gpurun_out/sanitize_rfa_summary.txt.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
out=gpurun_out/sanitize_rfa_summary.txt; : > $out
run() {   # tool, tag, pytest -k expression
  local tool=$1 tag=$2 expr=$3 log=gpurun_out/san_${1}_${2}.log
  timeout ${SAN_TIMEOUT:-400} compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_rfa_gpu.py -m gpu -q -x -p no:cacheprovider -k "$expr" > $log 2>&1
  local rc=$?
  {
    echo "== $tool / $tag  (pytest tests/test_rfa_gpu.py -k \"$expr\")  rc=$rc"
    echo "   $(grep -a -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
    echo "   pytest: $(grep -a -E '[0-9]+ (passed|failed)' $log | tail -1)"
    grep -a -E '^=========     (Invalid|Out-of-range|Misaligned|Race|Barrier|Uninitialized|Program hit|Error:|Warning:|Potential|Divergent|Hazard)' $log | sort | uniq -c | sort -rn | head -6 | sed 's/^/   /'
  } >> $out
}
run memcheck performer_tc  "test_performer_tcgen05_path_vs_oracle and dtype0"
run memcheck performer_simt "test_performer_core_vs_oracle and dtype0"
run memcheck ra            "test_ra_core_vs_oracle and (dtype0 or dtype1)"
run memcheck scatterbrain  "test_scatterbrain_core_vs_oracle and (dtype0 or dtype1) and not many"
run racecheck performer_tc "(test_performer_tcgen05_path_vs_oracle and dtype0 and 203) or (running_stabiliser and dtype0)"
run racecheck ra           "test_ra_core_vs_oracle and dtype1 and gather"
run racecheck scatterbrain "test_scatterbrain_core_vs_oracle and dtype1 and (1d_w16_mask_m64 or 2d_w8_d32)"
run synccheck performer_tc "(test_performer_tcgen05_path_vs_oracle and dtype0 and 203) or (running_stabiliser and dtype0)"
run synccheck scatterbrain "test_scatterbrain_core_vs_oracle and dtype1 and 1d_w16_mask_m64"
run memcheck ra_sample     "test_ra_sample_gumbel_max_vs_oracle and dtype0 and 203"
run racecheck ra_sample    "test_ra_sample_gumbel_max_vs_oracle and dtype0 and 203"
run synccheck ra_sample    "test_ra_sample_gumbel_max_vs_oracle and dtype0 and 203"
cat $out
