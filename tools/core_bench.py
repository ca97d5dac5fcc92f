"""Core-only timing + parity check of the fused kernel (development tool; bench.py is the contract).

    python tools/core_bench.py [batch] [iters]

Prints tokens/s of eva_forward (C ABI, q/k/v resident in HBM) and the relative L2 error against the
generic CUDA-core path (EVA_SM100_DISABLE_FUSED is read once per process, so the generic result comes
from a child process and is cached in gpurun_out/)."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from efficient_attention import _abi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mode = sys.argv[3] if len(sys.argv) > 3 else 'time'
dev = torch.device('cuda', 0)
layer = bench.build_layer(dev, torch.float16)
torch.manual_seed(1)
PB = 8   # parity batch
with torch.no_grad():
    if mode == 'ref':   # child: generic path on a small batch -> file
        x = torch.randn(PB, 28, 28, 192, device=dev, dtype=torch.float16)
        q, k, v, _ = layer._qkv_heads(x.reshape(PB, 784, 192))
        geom = _abi.eva_geometry(q, seq_shape=(28, 28), window=7, ext=0, chunk=4, chunk_ext=0)
        ada, bias = layer._adaptive(), layer._local_bias().float().contiguous()
        out, path = _abi.eva_forward(q, k, v, geom, ada, bias=bias, return_path=True)
        assert path != 1
        torch.save(out.float().cpu(), os.path.join(ROOT, 'gpurun_out', 'core_ref.pt'))
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    env = dict(os.environ, EVA_SM100_DISABLE_FUSED='1')
    subprocess.run([sys.executable, __file__, str(B), '1', 'ref'], env=env, check=True)
    ref = torch.load(os.path.join(ROOT, 'gpurun_out', 'core_ref.pt'))
    x = torch.randn(PB, 28, 28, 192, device=dev, dtype=torch.float16)
    q, k, v, _ = layer._qkv_heads(x.reshape(PB, 784, 192))
    geom = _abi.eva_geometry(q, seq_shape=(28, 28), window=7, ext=0, chunk=4, chunk_ext=0)
    ada, bias = layer._adaptive(), layer._local_bias().float().contiguous()
    out, path = _abi.eva_forward(q, k, v, geom, ada, bias=bias, return_path=True)
    torch.cuda.synchronize()
    assert path == 1, 'fused path not taken'
    err = float((out.float().cpu() - ref).norm() / ref.norm())
    print(f'parity (B={PB}) fused vs generic: rel-L2 {err:.3e}', flush=True)

    x = torch.randn(B, 28, 28, 192, device=dev, dtype=torch.float16)
    q, k, v, _ = layer._qkv_heads(x.reshape(B, 784, 192))
    geom = _abi.eva_geometry(q, seq_shape=(28, 28), window=7, ext=0, chunk=4, chunk_ext=0)
    for _ in range(5):
        _abi.eva_forward(q, k, v, geom, ada, bias=bias)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import threading
    import pynvml
    pynvml.nvmlInit()
    hdl = pynvml.nvmlDeviceGetHandleByIndex(0)
    clocks, stop = [], threading.Event()

    def sample():
        while not stop.is_set():
            clocks.append((pynvml.nvmlDeviceGetClockInfo(hdl, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hdl) / 1000))
            stop.wait(0.002)
    th = threading.Thread(target=sample)
    reps = max(1, int(0.25 / (0.00045 * iters)))      # ~0.25 s of back-to-back launches so that the sampler sees them
    th.start()
    e0.record()
    for _ in range(iters * reps):
        _abi.eva_forward(q, k, v, geom, ada, bias=bias)
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    ms = e0.elapsed_time(e1) / (iters * reps)
    cs = sorted(c for c, _ in clocks[len(clocks) // 3:])
    ps = sorted(w for _, w in clocks[len(clocks) // 3:])
    print(f'clocks: sm median {cs[len(cs) // 2]} MHz (min {cs[0]}, max {cs[-1]}), power median {ps[len(ps) // 2]:.0f} W, {len(cs)} samples')
    tps = B * 784 / (ms * 1e-3)
    print(f'core: B={B} {ms * 1e3:.1f} us/launch  {tps / 1e9:.3f} G tokens/s  ({tps * 1536 / 6469.3e9 * 100:.1f} % of HBM roofline)'
          f'  env: ' + ' '.join(f'{k_}={v_}' for k_, v_ in os.environ.items() if k_.startswith('EVA_SM100')), flush=True)
    assert err < 2e-3, err
