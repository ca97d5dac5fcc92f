"""eva_forward on geometries outside the fused kernels: time of the two stages (chunk statistics | window attention), fp16."""
import math
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def timed(fn, n=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    import bench
    bench.use_product_package()
    from efficient_attention import _abi
    from test_gpu_parity import _abi_ada, _rand_ada
    dev = torch.device('cuda', 0)
    cases = [('EVA 28x28, window 7, halo 3, chunk 4 (+halo 3), B=128', (28, 28), 7, 3, 4, False, 128, 3, False),
             ('EVA 14x14, window 7, halo 3, chunk 2 (+halo 3), B=512', (14, 14), 7, 3, 2, False, 512, 6, False),
             ('EVA 1-D N=1024, window 64, halo 32, chunk 16 (+halo 32), B=64, padding mask', (1024,), 64, 32, 16, False, 64, 8, True),
             ('EVA 28x28, window 14, no halo, chunk 4, B=128', (28, 28), 14, 0, 4, False, 128, 3, False),
             ('causal T=4096, window 128 + left halo 128, chunk 128, B=16', (4096,), 128, 128, 128, True, 16, 8, False)]
    for name, seq_shape, window, ext, chunk, causal, B, H, with_mask in cases:
        d = 64
        N = math.prod(seq_shape)
        g = torch.Generator().manual_seed(0)
        qkv = torch.randn(B, N, 3, H, d, generator=g).to(dev, torch.float16)
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        mask = None
        if with_mask:
            mask = torch.zeros(B, N, dtype=torch.bool, device=dev)
            mask[:, N - 100:] = True
        geometry = dict(seq_shape=seq_shape, window=window, ext=ext, chunk=chunk, chunk_ext=0 if causal else ext, causal=causal,
                        halo_left_only=causal, mask_queries=causal)
        geom = _abi.eva_geometry(q, **geometry)
        ada = _abi_ada(_rand_ada(d, g), dev, 1.0 if causal else 0.5)
        kb, bt = _abi.eva_chunk_stats(q, k, v, geom, ada, pad_mask=mask)
        t_stats = timed(lambda: _abi.eva_chunk_stats(q, k, v, geom, ada, pad_mask=mask))
        t_win = timed(lambda: _abi.eva_window_attention(q, k, v, geom, k_bar=kb, beta=bt, pad_mask=mask))
        out, path = _abi.eva_forward(q, k, v, geom, ada, pad_mask=mask, return_path=True)
        t_all = timed(lambda: _abi.eva_forward(q, k, v, geom, ada, pad_mask=mask))
        print(f'{name}: path {path}  statistics {t_stats:.3f} ms | window {t_win:.3f} ms | eva_forward {t_all:.3f} ms '
              f'({B * N / (t_all * 1e-3) / 1e6:.0f} M tokens/s)')


if __name__ == '__main__':
    main()
