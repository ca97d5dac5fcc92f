#!/bin/bash
# round-end evidence for the headline kernel: GPU tests, smoke, bench line, ncu launch list + full capture (text exports)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eva_fused -s 6 -c 1 -f -o gpurun_out/fused_full \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/fused_full.ncu-rep --page raw --csv > gpurun_out/ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/fused_full.ncu-rep --page details --csv > gpurun_out/ncu_details.csv 2>/dev/null
rm -f gpurun_out/fused_full.ncu-rep
cut -c1-1200 gpurun_out/bench.json; echo; cut -c1-600 gpurun_out/bench_ref.json
