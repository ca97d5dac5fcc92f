#!/bin/bash
# calibration: core tokens/s and phase traces at 2 and 1 CTAs per SM
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python tools/core_bench.py 1024 20 2>&1 | tail -3 | tee gpurun_out/core2.log
EVA_SM100_CTAS_PER_SM=1 timeout 300 python tools/core_bench.py 1024 20 2>&1 | tail -2 | tee gpurun_out/core1.log
timeout 200 python tools/trace_dump.py 1024 > gpurun_out/trace2.log 2>&1
EVA_SM100_CTAS_PER_SM=1 timeout 200 python tools/trace_dump.py 1024 > gpurun_out/trace1.log 2>&1
echo done
