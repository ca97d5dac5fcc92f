"""Which part of bench.py's timing harness costs the headline kernel ~8 % against a bare launch loop?  (development tool)"""
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
    dev = torch.device('cuda', 0)
    B, K = 1024, 20
    layer = bench.build_layer(dev, torch.float16)
    torch.manual_seed(1)
    x = torch.randn(B, 28, 28, 192, device=dev, dtype=torch.float16)
    with torch.no_grad():
        q, k, v, _ = layer._qkv_heads(x.reshape(B, 784, 192))
        geom = _abi.eva_geometry(q, seq_shape=(28, 28), window=7, ext=0, chunk=4, chunk_ext=0)
        ada, bias = layer._adaptive(), layer._local_bias().float().contiguous()
        core = lambda: _abi.eva_forward(q, k, v, geom, ada, bias=bias, return_path=True)

        def run(per_launch_events, sampler, idle=1.0):
            time.sleep(idle)                       # let the board cool down / leave the power cap
            clk = bench.ClockSampler(0) if sampler else None
            if clk:
                clk.__enter__()
            for _ in range(5):
                core()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            e0.record()
            for a, b in pairs:
                if per_launch_events:
                    a.record()
                core()
                if per_launch_events:
                    b.record()
            e1.record()
            torch.cuda.synchronize()
            if clk:
                clk.__exit__(None, None, None)
            tot = e0.elapsed_time(e1) / K
            per = sum(a.elapsed_time(b) for a, b in pairs) / K if per_launch_events else float('nan')
            return tot, per

        for rep in range(2):
            for ple in (False, True):
                for smp in (False, True):
                    tot, per = run(ple, smp)
                    print(f'rep {rep} per-launch events {ple!s:5} NVML sampler {smp!s:5}: {tot * 1e3:6.1f} us/step (mean of per-launch pairs {per * 1e3:6.1f})', flush=True)


if __name__ == '__main__':
    main()
