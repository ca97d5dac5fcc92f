"""Diagnostics for the fused tcgen05/TMA EVA kernel: compare eva_forward (fused path) with the generic
two-stage kernels and the CPU oracle on identical fp16 inputs, and localise any mismatch.
Run on the GPU box:  python tools/fused_diag.py [B] > gpurun_out/fused_diag.log"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200'))
sys.path.insert(0, ROOT)
from efficient_attention import _abi  # noqa: E402
from oracle import eva_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    grid = int(sys.argv[2]) if len(sys.argv) > 2 else 28
    dtype = torch.float16
    H, d, w = 3, 64, 7
    N = grid * grid
    chunk = int(math.sqrt(N // 49))
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.2).to(dtype)
    bias = 0.5 * torch.randn(H, 49, 49, generator=g)
    wq, wk = torch.randn(d, d, generator=g) / 8, torch.randn(d, d, generator=g) / 8
    bq, bk = 0.1 * torch.randn(d, generator=g), 0.1 * torch.randn(d, generator=g)
    gq, gk = 1 + 0.1 * torch.randn(d, generator=g), 1 + 0.1 * torch.randn(d, generator=g)
    eq, ek = 0.1 * torch.randn(d, generator=g), 0.1 * torch.randn(d, generator=g)
    q64, k64, v64 = (qkv[:, :, i].permute(0, 2, 1, 3).double() for i in range(3))
    want, kbar_w, beta_w = O.eva_core(q64, k64, v64, seq_shape=(grid, grid), window=w, ext=0, chunk=chunk, chunk_ext=0,
                                      wq=wq.double(), bq=bq.double(), gq=gq.double(), betq=eq.double(),
                                      wk=wk.double(), bk=bk.double(), gk=gk.double(), betk=ek.double(),
                                      mu_coeff=0.5, bias=bias.double(), return_stats=True)
    want = want.permute(0, 2, 1, 3).reshape(B, N, H * d)
    qd = qkv.to(dev)
    q, k, v = qd[:, :, 0], qd[:, :, 1], qd[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(grid, grid), window=w, ext=0, chunk=chunk, chunk_ext=0)
    mv = lambda t: t.to(dev)
    ada = _abi.adaptive(mv(wq), mv(bq), mv(gq), mv(eq), mv(wk), mv(bk), mv(gk), mv(ek), mu_coeff=0.5)
    kb, bt = _abi.eva_chunk_stats(q, k, v, geom, ada)
    gen = _abi.eva_window_attention(q, k, v, geom, k_bar=kb, beta=bt, bias=bias.to(dev))
    torch.cuda.synchronize()
    print(f'generic vs oracle: {rel(gen.cpu(), want):.3e}')
    out, path = _abi.eva_forward(q, k, v, geom, ada, bias=bias.to(dev), return_path=True)
    torch.cuda.synchronize()
    print(f'path taken: {path} (1 = fused)')
    o = out.cpu().double()
    print(f'fused vs oracle : {rel(o, want):.3e}   nan={int(torch.isnan(o).sum())} inf={int(torch.isinf(o).sum())}')
    print(f'fused vs generic: {rel(o, gen.cpu()):.3e}')
    # localise: [B, gh/w, w, gw/w, w, H, d]
    nw = grid // w
    err = (o - want).reshape(B, nw, w, nw, w, H, d)
    ref = want.reshape(B, nw, w, nw, w, H, d)
    def by(dims, label):
        keep = [i for i in range(7) if i not in dims]
        e = err.pow(2).sum(keep).sqrt() / ref.pow(2).sum(keep).sqrt().clamp(min=1e-30)
        print(label, [f'{x:.1e}' for x in e.flatten().tolist()][:64])
    by([0], 'per batch      ')
    by([5], 'per head       ')
    by([1, 3], 'per window     ')
    by([2, 4], 'per row in win ')
    e = err.reshape(-1, d).pow(2).sum(0).sqrt() / ref.reshape(-1, d).pow(2).sum(0).sqrt()
    print('per feature    ', [f'{x:.1e}' for x in e.tolist()])
    # hypotheses
    local_only = O.local_core(q64, k64, v64, seq_shape=(grid, grid), window=w, ext=0, bias=bias.double())
    print(f'fused vs local-only attention: {rel(o, local_only.permute(0, 2, 1, 3).reshape(B, N, H * d)):.3e}')
    nb, _, _ = O.eva_core(q64, k64, v64, seq_shape=(grid, grid), window=w, ext=0, chunk=chunk, chunk_ext=0,
                          wq=wq.double(), bq=bq.double(), gq=gq.double(), betq=eq.double(),
                          wk=wk.double(), bk=bk.double(), gk=gk.double(), betk=ek.double(),
                          mu_coeff=0.5, bias=None, return_stats=True)
    print(f'fused vs no-bias oracle      : {rel(o, nb.permute(0, 2, 1, 3).reshape(B, N, H * d)):.3e}')
    print('sample out[0,0,:8]  ', [f'{x:.4f}' for x in o[0, 0, :8].tolist()])
    print('sample want[0,0,:8] ', [f'{x:.4f}' for x in want[0, 0, :8].tolist()])
    print('sample out[0,57,:8] ', [f'{x:.4f}' for x in o[0, 57, :8].tolist()])
    print('sample want[0,57,:8]', [f'{x:.4f}' for x in want[0, 57, :8].tolist()])
    # bf16 path
    qb = qkv.to(torch.bfloat16).to(dev)
    outb, pathb = _abi.eva_forward(qb[:, :, 0], qb[:, :, 1], qb[:, :, 2], geom.__class__(*[getattr(geom, f) for f, _ in geom._fields_][:-1], _abi.EVA_BF16),
                                   ada, bias=bias.to(dev), return_path=True)
    qb64 = [qb[:, :, i].permute(0, 2, 1, 3).double().cpu() for i in range(3)]
    wantb = O.eva_core(*qb64, seq_shape=(grid, grid), window=w, ext=0, chunk=chunk, chunk_ext=0,
                       wq=wq.double(), bq=bq.double(), gq=gq.double(), betq=eq.double(),
                       wk=wk.double(), bk=bk.double(), gk=gk.double(), betk=ek.double(), mu_coeff=0.5, bias=bias.double())
    print(f'bf16 path {pathb}: fused vs oracle {rel(outb.cpu(), wantb.permute(0, 2, 1, 3).reshape(B, N, H * d)):.3e}')


if __name__ == '__main__':
    main()
