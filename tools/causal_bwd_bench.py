"""Causal EVA (c5: T = 4096, h = 8, window = chunk = 256, batch 16, fp16) forward + backward of the core: `eva_backward` kernels vs
autograd through the PyTorch recomputation (float32, and under fp16 autocast)  (development tool)."""
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
    from efficient_attention import _recompute
    from test_gpu_parity import _rand_ada
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device('cuda', 0)
    H, d, N, w = 8, 64, 4096, 256
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(B, N, 3, H, d, generator=g).to(dev, torch.float16)
    ada = {k_: v_.to(dev).requires_grad_(True) for k_, v_ in _rand_ada(d, g).items()}
    names = ('wq', 'bq', 'gq', 'betq', 'wk', 'bk', 'gk', 'betk')
    noise = torch.randn(B, H, N // w, d, generator=g).to(dev)
    bias = (0.1 * torch.randn(1, w, 2 * w, generator=g)).to(dev)
    ext = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # 0: overlap_window=False (the c5 configuration); 256: overlapping windows
    bias = bias[:, :, :w + ext].contiguous() if os.environ.get('NO_BIAS') != '1' else None
    geometry = dict(seq_shape=(N,), window=w, ext=ext, chunk=w, chunk_ext=0, causal=True, halo_left_only=True, mask_queries=True)
    wgt = torch.randn(B, N, H * d, generator=g).to(dev, torch.float16)
    for impl in ('cuda', 'torch'):
        prev = _recompute.set_backward_impl(impl)
        try:
            def step():
                x = qkv.detach().requires_grad_(True)
                out = _recompute.eva_core(x[:, :, 0], x[:, :, 1], x[:, :, 2], geometry=geometry, mu_coeff=1.0,
                                          params=[ada[n] for n in names], noise=noise, bias=bias)
                (out * wgt).sum().backward()
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                step()
                b_.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b_))
            print(f'backward impl {impl:5s}: forward + backward of the core {statistics.median(ts):.2f} ms, peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB')
            if impl == 'cuda':
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    step()
                    torch.cuda.synchronize()
                rows = sorted(((e.device_time_total, e.count, e.key) for e in prof.key_averages() if e.device_time_total > 0), reverse=True)
                for t, n, name in rows[:6]:
                    print(f'    {t / 1e3:8.3f} ms x{n:<3d} {name[:100]}')
        finally:
            _recompute.set_backward_impl(prev)


if __name__ == '__main__':
    main()
