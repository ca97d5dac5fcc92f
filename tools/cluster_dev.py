"""Development check of the cluster-resident fused kernel (path 3) against the oracle + timing vs the streamed kernel (path 1).
usage: python tools/cluster_dev.py [B] [reps]"""
import ctypes
import math
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200'))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from efficient_attention import _abi
    from oracle import eva_oracle as O
    from test_gpu_parity import _abi_ada, _rand_ada
    dev = torch.device('cuda', 0)
    if os.environ.get('EVA_SM100_DISABLE_CLUSTER') != '1':
        _abi.load().eva_debug_set_cluster_mode(ctypes.c_int(1))
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    H, d, grid, chunk = 3, 64, 28, 4
    N = grid * grid
    g = torch.Generator().manual_seed(3)
    # ---- correctness on a small batch, several items per cluster round ------------------------------------
    for Bs, with_bias, with_noise in ((2, True, False), (5, False, True), (60, True, True)):
        qkv = (torch.randn(Bs, N, 3, H, d, generator=g) * 1.1).half()
        bias = 0.5 * torch.randn(H, 49, 49, generator=g) if with_bias else None
        noise = torch.randn(Bs, H, 49, d, generator=g) if with_noise else None
        ada = _rand_ada(d, g)
        q64, k64, v64 = (qkv[:, :, i].permute(0, 2, 1, 3).double() for i in range(3))
        want = O.eva_core(q64, k64, v64, seq_shape=(grid, grid), window=7, ext=0, chunk=chunk, chunk_ext=0,
                          **{k_: v_.double() for k_, v_ in ada.items()}, mu_coeff=0.5,
                          noise=noise.double() if with_noise else None, bias=bias.double() if with_bias else None)
        want = want.permute(0, 2, 1, 3).reshape(Bs, N, H * d)
        qd = qkv.to(dev)
        geom = _abi.eva_geometry(qd[:, :, 0], seq_shape=(grid, grid), window=7, ext=0, chunk=chunk, chunk_ext=0)
        out, path = _abi.eva_forward(qd[:, :, 0], qd[:, :, 1], qd[:, :, 2], geom, _abi_ada(ada, dev, 0.5),
                                     noise=noise.to(dev) if with_noise else None, bias=bias.to(dev) if with_bias else None,
                                     return_path=True)
        torch.cuda.synchronize()
        o = out.cpu().double()
        err = float((o - want).norm() / want.norm())
        per_item = ((o - want).view(Bs, N, H, d).pow(2).sum((1, 3)).sqrt() / want.view(Bs, N, H, d).pow(2).sum((1, 3)).sqrt())
        # where is it wrong? per window-pair / half
        e_tok = (o - want).view(Bs, grid, grid, H * d).pow(2).sum(-1).sqrt().mean(0)
        print(f'B={Bs} bias={with_bias} noise={with_noise}: path {path} rel-L2 {err:.3e} worst item {float(per_item.max()):.3e} nan {bool(torch.isnan(o).any())}')
        if err > 1e-3:
            torch.set_printoptions(linewidth=250, precision=2, sci_mode=False)
            print((e_tok / want.view(Bs, grid, grid, H * d).pow(2).sum(-1).sqrt().mean(0)))
    # ---- timing ---------------------------------------------------------------------------------------------
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.1).half().to(dev)
    bias = (0.5 * torch.randn(H, 49, 49, generator=g)).to(dev)
    ada = _abi_ada(_rand_ada(d, g), dev, 0.5)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(grid, grid), window=7, ext=0, chunk=chunk, chunk_ext=0)
    for _ in range(10):
        out, path = _abi.eva_forward(q, k, v, geom, ada, bias=bias, return_path=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _abi.eva_forward(q, k, v, geom, ada, bias=bias)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = B * N * 1536 / (ms * 1e-3) / 1e9
    print(f'B={B} path {path}: {ms * 1e3:.1f} us per launch, {B * N / (ms * 1e-3) / 1e9:.3f} G tokens/s, {gbs:.0f} GB/s = {gbs / 6545.9 * 100:.1f} % of 6545.9')
    if os.environ.get('EVA_SM100_TRACE') == '1':
        lib = _abi.load()
        buf = (ctypes.c_ulonglong * 20)()
        lib.eva_debug_read_cluster_prof(buf)
        n = max(1, buf[16])
        names = ['start->q copied,|k|^2', '->pool MMAs done', '->means written', '->Linear done', '->LN rows written', '->phi-logits ready',
                 '->logits exchanged', '->beta rows written', '->rows sent', '->pair0 S ready', '->pair0 P written', '->pair0 O ready',
                 '->pair1 S ready (incl. pair0 epilogue)', '->pair1 P written', '->pair1 O ready', '->item end (pair1 epilogue)']
        tot = 0
        for i, nm in enumerate(names):
            print(f'   {nm:45s} {buf[i] / n:8.0f} cyc')
            tot += buf[i] / n
        print(f'   item total {tot:.0f} cyc; gap between items {buf[17] / max(1, n - 1):.0f} cyc ({n} items of cluster 0 rank 0, warpgroup 0 thread 0)')


if __name__ == '__main__':
    main()
