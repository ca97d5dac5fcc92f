#!/bin/bash
# compute-sanitizer memcheck over the three tcgen05 kernels (small shapes)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for t in "test_fused_kernel_many_items_per_cta_fp16" "test_causal_tcgen05_many_windows_per_cta_fp16" "test_lara_tcgen05_core_many_items_fp16"; do
  timeout 500 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "$t" > gpurun_out/san_$t.log 2>&1
  echo "== $t: $(grep -a -c 'Invalid\|Out-of-range\|misaligned' gpurun_out/san_$t.log) findings; $(grep -a 'ERROR SUMMARY' gpurun_out/san_$t.log | tail -1); $(grep -a 'passed\|failed' gpurun_out/san_$t.log | tail -1)"
done
