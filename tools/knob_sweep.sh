#!/bin/bash
# sweep runtime tuning knobs of the fused kernel; prints tokens/s per setting
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for cfg in "7 1" "0 1" "2 1" "4 1" "7 0" "0 0"; do
  set -- $cfg
  v=$(EVA_SM100_PREFETCH_ROWS=$1 EVA_SM100_PREFETCH_V=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g' % d['value'])")
  echo "prefetch_rows=$1 prefetch_v=$2 -> $v tok/s" | tee -a gpurun_out/knobs.log
done
