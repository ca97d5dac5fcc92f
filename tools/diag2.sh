#!/bin/bash
# A/B on one box: baseline library (lib/libeva_sm100_base.so, if present) vs the current build; then the phase trace
# (+ GPU tests when asked: diag2.sh tests).  EXTRA="VAR=val VAR2=val2" adds one more timed configuration of the new build.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
BASE=efficient-attention_b200/lib/libeva_sm100_base.so
for i in 1 2; do
  if [ -f $BASE ]; then echo "base:"; EVA_SM100_LIB=$PWD/$BASE timeout 300 python tools/core_bench.py 1024 20 2>&1 | tail -1; fi
  echo "new:"; timeout 300 python tools/core_bench.py 1024 20 2>&1 | tail -3
  if [ -n "$EXTRA" ]; then echo "new + $EXTRA:"; env $EXTRA timeout 300 python tools/core_bench.py 1024 20 2>&1 | tail -1; fi
done | tee gpurun_out/core2.log
timeout 200 python tools/trace_dump.py 1024 > gpurun_out/trace2.log 2>&1
if [ "$1" = "tests" ]; then timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/pytest.log; fi
echo done
