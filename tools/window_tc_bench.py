"""`eva_window_attention` on geometries outside the fused kernels: tcgen05 generic kernel (eva_window_tc_sm100.cu) vs the CUDA-core
window_attn_kernel, fp16 (development tool)."""
import math
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import bench
    bench.use_product_package()
    from efficient_attention import _abi
    from test_gpu_parity import _abi_ada, _rand_ada
    dev = torch.device('cuda', 0)
    lib = _abi.load()
    cases = [('EVA 28x28, window 7, halo 3 (overlap_window), 49 chunks, B=128', (28, 28), 7, 3, 4, False, 128, 3),
             ('EVA 1-D N=1024, window 64, halo 32, chunk 16, B=64', (1024,), 64, 32, 16, False, 64, 8),
             ('local 14x14 window 7, no chunks, B=512', (14, 14), 7, 0, 0, False, 512, 6),
             ('softmax N=196 (DeiT-small-p16 baseline), B=512', (196,), 196, 0, 0, False, 512, 6),
             ('softmax N=784 (DeiT-tiny-p8 baseline), B=128', (784,), 784, 0, 0, False, 128, 3),
             ('causal T=4096, window 128 + left halo 128, chunk 128, B=16', (4096,), 128, 128, 128, True, 16, 8)]
    for name, seq_shape, window, ext, chunk, causal, B, H in cases:
        d = 64
        N = math.prod(seq_shape)
        g = torch.Generator().manual_seed(0)
        qkv = torch.randn(B, N, 3, H, d, generator=g).to(dev, torch.float16)
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        geometry = dict(seq_shape=seq_shape, window=window, ext=ext, chunk=chunk, chunk_ext=0 if causal else ext, causal=causal,
                        halo_left_only=causal, mask_queries=causal, mask_is_neg_inf=(chunk == 0 and window == N))
        geom = _abi.eva_geometry(q, **geometry)
        stats = {}
        if chunk:
            kb, bt = _abi.eva_chunk_stats(q, k, v, geom, _abi_ada(_rand_ada(d, g), dev, 1.0 if causal else 0.5))
            stats = dict(k_bar=kb, beta=bt)
        res = []
        for mode in (1, 0):
            lib.eva_debug_set_window_tc(mode)
            for _ in range(3):
                _abi.eva_window_attention(q, k, v, geom, **stats)
            torch.cuda.synchronize()
            ts = []
            for _ in range(8):
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _abi.eva_window_attention(q, k, v, geom, **stats)
                b_.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b_))
            res.append(statistics.median(ts))
        lib.eva_debug_set_window_tc(-1)
        gbs = B * N * H * d * 2 * 4 / (res[0] * 1e-3) / 1e9
        print(f'{name}: tcgen05 {res[0]:.3f} ms ({gbs:.0f} GB/s of q,k,v,out) | CUDA cores {res[1]:.3f} ms | x{res[1] / res[0]:.1f}')


if __name__ == '__main__':
    main()
