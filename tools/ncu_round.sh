#!/bin/bash
# ncu captures of the fused kernel: launch list of the bench + one full-set capture.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B=${1:-512}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --batch $B --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "launch-list exit $?" >> gpurun_out/ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eva_fused -s 2 -c 1 -f -o gpurun_out/fused_full \
  python bench.py --steps 2 --warmup 3 --batch $B --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "full exit $?" >> gpurun_out/ncu_full.log
tail -3 gpurun_out/ncu_bench.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
