"""Does the headline kernel's launch time depend on how long the GPU has been busy?  Times blocks of 20 launches back to back for
~1.5 s and prints the per-block average with the SM clock / power NVML reports at that moment (development tool)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200'))
import bench  # noqa: E402


def main():
    from efficient_attention import _abi
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device('cuda', 0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    layer = bench.build_layer(dev, torch.float16)
    torch.manual_seed(1)
    x = torch.randn(B, 28, 28, 192, device=dev, dtype=torch.float16)
    with torch.no_grad():
        q, k, v, _ = layer._qkv_heads(x.reshape(B, 784, 192))
        geom = _abi.eva_geometry(q, seq_shape=(28, 28), window=7, ext=0, chunk=4, chunk_ext=0)
        ada, bias = layer._adaptive(), layer._local_bias().float().contiguous()
        for _ in range(3):
            _abi.eva_forward(q, k, v, geom, ada, bias=bias)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        blk = 0
        while time.perf_counter() - t0 < 1.5:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                _abi.eva_forward(q, k, v, geom, ada, bias=bias)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            if blk < 12 or blk % 10 == 0:
                print(f't={time.perf_counter() - t0:6.3f}s block {blk:3d}: {ms * 1e3:6.1f} us/launch  sm {pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)} MHz '
                      f'mem {pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM)} MHz  {pynvml.nvmlDeviceGetPowerUsage(h) / 1000:.0f} W  '
                      f'reasons {pynvml.nvmlDeviceGetCurrentClocksEventReasons(h):#x}')
            blk += 1


if __name__ == '__main__':
    main()
