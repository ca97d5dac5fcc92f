#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 200 python tools/lara_bench.py 64 > gpurun_out/lara_small.log 2>&1; grep -a "eva fused\|timeout\|rror\|parity" gpurun_out/lara_small.log | head -6 | cut -c1-300
timeout 400 compute-sanitizer --tool memcheck --print-limit 3 python tools/c4_split.py > gpurun_out/sanitizer.log 2>&1; grep -a -A6 "=========" gpurun_out/sanitizer.log | head -40 | cut -c1-220
