#!/bin/bash
# One GPU-box session: fused-kernel diagnostics, parity tests, smoke, bench, phase trace. Everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 120 python tools/fused_diag.py 2 28 > gpurun_out/fused_diag.log 2>&1
rc=$?; echo "diag exit $rc" >> gpurun_out/fused_diag.log
timeout 120 python tools/fused_diag.py 3 14 > gpurun_out/fused_diag14.log 2>&1
echo "diag14 exit $?" >> gpurun_out/fused_diag14.log
if [ $rc -ne 0 ]; then export EVA_SM100_DISABLE_FUSED=1; echo "FUSED DISABLED for the rest of this run" >> gpurun_out/fused_diag.log; fi
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
if [ $rc -eq 0 ]; then timeout 200 python tools/trace_dump.py 1024 > gpurun_out/trace.log 2>&1; echo "trace exit $?" >> gpurun_out/trace.log; fi
grep -v "^per \|^sample" gpurun_out/fused_diag.log | tail -12; tail -3 gpurun_out/fused_diag14.log; tail -5 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
