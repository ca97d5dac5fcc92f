#!/bin/bash
# per-instruction stall samples of the fused LARA kernel (source page of one --set full capture), reduced to the hottest source lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lara_core -s 1 -c 1 -f -o gpurun_out/laracore python tools/c4_split.py > /dev/null 2>&1
ncu -i gpurun_out/laracore.ncu-rep --page source --csv > gpurun_out/lara_source.csv 2>/dev/null
rm -f gpurun_out/laracore.ncu-rep
head -c 1500 gpurun_out/lara_source.csv | head -3
wc -l gpurun_out/lara_source.csv
