"""Times `eva_backward` (libeva_sm100) on the c3 geometry: tcgen05 window kernel vs CUDA-core kernel, with / without the bias-table
gradient (development tool).

    python tools/bwd_bench.py [batch]
"""
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
    from test_gpu_parity import _rand_ada
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    dev = torch.device('cuda', 0)
    lib = _abi.load()
    H, d, N = 3, 64, 784
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(B, N, 3, H, d, generator=g).to(dev, torch.float16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    ada_t = {k_: v_.to(dev) for k_, v_ in _rand_ada(d, g).items()}
    ada = _abi.adaptive(*[ada_t[n] for n in ('wq', 'bq', 'gq', 'betq', 'wk', 'bk', 'gk', 'betk')], mu_coeff=0.5)
    bias = (0.5 * torch.randn(H, 49, 49, generator=g)).to(dev)
    noise = torch.randn(B, H, 49, d, generator=g).to(dev)
    geom = _abi.eva_geometry(q, seq_shape=(28, 28), window=7, ext=0, chunk=4, chunk_ext=0)
    out = _abi.eva_forward(q, k, v, geom, ada, noise=noise, bias=bias)
    gout = torch.randn_like(out)
    for mode, name in ((1, 'tcgen05'), (0, 'cuda cores')):
        lib.eva_debug_set_bwd_tc(mode)
        for with_bias, want in ((True, True), (True, False), (False, False)):
            def run():
                return _abi.eva_backward(q, k, v, geom, ada, out, gout, noise=noise, bias=bias if with_bias else None, want_bias_grad=want)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run()
                b_.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b_))
            print(f'{name:10s} bias={with_bias!s:5s} bias_grad={want!s:5s}: eva_backward (all kernels + zero fills) {statistics.median(ts):.3f} ms')
    lib.eva_debug_set_bwd_tc(-1)


if __name__ == '__main__':
    main()
