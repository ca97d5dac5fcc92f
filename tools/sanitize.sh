#!/bin/bash
# compute-sanitizer memcheck over the tcgen05 kernels (small shapes)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
i=0
for t in "test_fused_kernel_many_items_per_cta_fp16 and 14-2" "test_fused_kernel_many_items_per_cta_fp16 and 28-4" "test_fused_kernel_variants_fp16 and 14-2 and default and True-True" "test_causal_tcgen05_many_windows_per_cta_fp16" "test_causal_time_major_views_fp16" "test_lara_tcgen05_core_many_items_fp16"; do
  i=$((i + 1)); log=gpurun_out/san_$i.log
  timeout 500 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "$t" > $log 2>&1
  echo "== $t: $(grep -a -c 'Invalid\|Out-of-range\|misaligned' $log) findings; $(grep -a 'ERROR SUMMARY' $log | tail -1); $(grep -a 'passed\|failed' $log | tail -1)"
done
