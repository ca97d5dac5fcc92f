"""eva_backward on geometries outside the fused kernels: kernel times by name (torch.profiler), fp16."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import bench
    bench.use_product_package()
    from efficient_attention import _recompute
    from test_gpu_parity import _rand_ada
    from torch.profiler import ProfilerActivity, profile
    dev = torch.device('cuda', 0)
    cases = [('EVA 28x28, window 7, halo 3, chunk 4 (+halo 3), B=128', (28, 28), 7, 3, 4, False, 128, 3),
             ('EVA 1-D N=1024, window 64, halo 32, chunk 16 (+halo 32), B=64', (1024,), 64, 32, 16, False, 64, 8),
             ('causal T=4096, window = chunk = 256, B=16', (4096,), 256, 0, 256, True, 16, 8)]
    names = ('wq', 'bq', 'gq', 'betq', 'wk', 'bk', 'gk', 'betk')
    for name, seq_shape, window, ext, chunk, causal, B, H in cases:
        d = 64
        N = math.prod(seq_shape)
        g = torch.Generator().manual_seed(0)
        qkv = torch.randn(B, N, 3, H, d, generator=g).to(dev, torch.float16)
        ada = {k_: v_.to(dev).requires_grad_(True) for k_, v_ in _rand_ada(d, g).items()}
        geometry = dict(seq_shape=seq_shape, window=window, ext=ext, chunk=chunk, chunk_ext=0 if causal else ext, causal=causal,
                        halo_left_only=causal, mask_queries=causal)
        noise = torch.randn(B, H, _recompute.num_chunks_of(seq_shape, chunk), d, generator=g).to(dev)
        w = torch.randn(B, N, H * d, generator=g).to(dev, torch.float16)

        def step():
            x = qkv.detach().requires_grad_(True)
            out = _recompute.eva_core(x[:, :, 0], x[:, :, 1], x[:, :, 2], geometry=geometry, mu_coeff=1.0 if causal else 0.5,
                                      params=[ada[n] for n in names], noise=noise, packed=x)
            (out * w).sum().backward()
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        rows = sorted(((e.device_time_total, e.count, e.key) for e in prof.key_averages() if e.device_time_total > 0), reverse=True)
        print(name)
        for t, n, kname in rows[:6]:
            print(f'    {t / 1e3:8.3f} ms x{n:<3d} {kname[:90]}')


if __name__ == '__main__':
    main()
