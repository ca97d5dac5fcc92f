#!/bin/bash
# ncu --set full captures of the causal window kernel, the CTA-per-chunk statistics kernel and the fused LARA kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:eva_causal_window -s 2 -c 1 -f -o gpurun_out/causal python tools/c5_split.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:chunk_stats_cta -s 2 -c 1 -f -o gpurun_out/cstats python tools/c5_split.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:lara_core -s 1 -c 1 -f -o gpurun_out/laracore python tools/c4_split.py > /dev/null 2>&1
for n in causal cstats laracore; do
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$n.csv 2>/dev/null; rm -f gpurun_out/$n.ncu-rep
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/ncu_raw_$n.csv")))
d=dict(zip(rows[0],rows[2])); u=dict(zip(rows[0],rows[1]))
ks=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","lts__t_sector_hit_rate.pct","launch__registers_per_thread","sm__warps_active.avg.pct_of_peak_sustained_active"]
print("$n:", "; ".join(f"{k.split('.')[0]}={d.get(k)} {u.get(k)}" for k in ks))
PY
done
