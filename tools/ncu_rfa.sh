#!/bin/bash
# ncu --set full of the tcgen05 random-feature kernels on the c3 geometry (one launch each), raw CSV -> gpurun_out/
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
cat > /tmp/rfa_one.py <<'PY'
import sys, torch
sys.path.insert(0, 'efficient-attention_b200')
from efficient_attention import _abi
dev = torch.device('cuda', 0)
B = 256
torch.manual_seed(0)
qkv = torch.randn(B, 784, 3, 3, 64, device=dev, dtype=torch.float16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
proj = torch.randn(3, 64, 64, device=dev)
bias = 0.5 * torch.randn(3, 49, 49, device=dev)
for _ in range(2):
    _abi.rfa_forward(q, k, v, method='favorp', proj=proj)
    _abi.scatterbrain_forward(q, k, v, seq_shape=(28, 28), window=7, proj=proj, bias=bias)
torch.cuda.synchronize()
PY
for kern in rfa_favorp_tc_kernel sb_window_tc_kernel; do
  timeout 600 ncu --set full --clock-control none -k regex:$kern -s 1 -c 1 --csv --page raw --log-file gpurun_out/ncu_raw_${kern}_B256.csv python /tmp/rfa_one.py > gpurun_out/ncu_${kern}.log 2>&1
  echo "$kern rc=$?"
done
