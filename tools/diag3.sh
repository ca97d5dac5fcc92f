#!/bin/bash
# first-failure diagnostics of the current build (small batch, full stderr)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 120 python tools/core_bench.py 8 2 > gpurun_out/core_small.log 2>&1; grep -a "eva fused\|timeout\|rror" gpurun_out/core_small.log | head -5 | cut -c1-300
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tools/core_bench.py 8 1 time > gpurun_out/sanitizer.log 2>&1; grep -a -A12 "=========" gpurun_out/sanitizer.log | head -60 | cut -c1-250
