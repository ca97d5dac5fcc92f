"""Training path (`-m gpu`): kernel forward + backward by recomputation (`efficient_attention/_recompute.py`).

* the float32 PyTorch recomputation equals the kernels' forward (it is what autograd differentiates);
* gradients of the drop-in modules (w.r.t. the input and every parameter) equal autograd through the float64 CPU oracle on the
  reference-generated training fixtures (identical noise draw);
* an fp16-autocast training step through the fused tcgen05 path produces finite gradients;
* attention-probability dropout of the causal layer (causal_eva.py:778);
* two-rank DDP step with the NCCL gradient all-reduce (vit/main.py:286-288) -- needs two GPUs, skipped otherwise.
"""
import math
import os
import subprocess
import sys

import pytest
import torch

from conftest import load_golden, rel_l2
from helpers import build_module
from oracle import eva_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dev():
    return torch.device('cuda', 0)


@pytest.mark.parametrize('seq_shape,window,ext,chunk,causal,with_mask', [
    ((14, 14), 7, 0, 2, False, False), ((28, 28), 7, 0, 4, False, False), ((14, 14), 7, 3, 2, False, False),
    ((96,), 16, 8, 12, False, True), ((128,), 32, 0, 16, True, True), ((96,), 16, 16, 8, True, False)])
def test_recomputation_equals_the_kernels_fp32(seq_shape, window, ext, chunk, causal, with_mask):
    """`eva_core_torch` (what the backward differentiates) against `eva_forward` of the library on identical float32 inputs."""
    from efficient_attention import _abi, _recompute
    from test_gpu_parity import _rand_ada
    dev = _dev()
    B, H, d = 2, 2, 64
    N = math.prod(seq_shape)
    g = torch.Generator().manual_seed(N + window + ext)
    qkv = torch.randn(B, N, 3, H, d, generator=g).to(dev)
    two_d = len(seq_shape) == 2
    L = window * window if two_d else window
    J = (window + 2 * ext) ** 2 if two_d else window + (ext if causal else 2 * ext)
    bias = (0.5 * torch.randn(H, L, J, generator=g)).to(dev)
    ada = {k_: (v_.to(dev) if v_ is not None else None) for k_, v_ in _rand_ada(d, g).items()}
    mask = None
    if with_mask:
        mask = torch.zeros(B, N, dtype=torch.bool, device=dev)
        mask[1, N - 9:] = True
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    chunk_ext = 0 if causal else ext
    n_chunks = _recompute.num_chunks_of(seq_shape, chunk)
    noise = torch.randn(B, H, n_chunks, d, generator=g).to(dev)
    coeff = 1.0 if causal else 0.5
    geometry = dict(seq_shape=seq_shape, window=window, ext=ext, chunk=chunk, chunk_ext=chunk_ext, causal=causal,
                    halo_left_only=causal, mask_queries=causal)
    geom = _abi.eva_geometry(q, **geometry)
    out = _abi.eva_forward(q, k, v, geom, _abi.adaptive(ada['wq'], ada['bq'], ada['gq'], ada['betq'], ada['wk'], ada['bk'], ada['gk'],
                                                        ada['betk'], mu_coeff=coeff), pad_mask=mask, noise=noise, bias=bias)
    ref = _recompute.eva_core_torch(q, k, v, seq_shape=seq_shape, window=window, ext=ext, chunk=chunk, chunk_ext=chunk_ext, **ada,
                                    mu_coeff=coeff, pad_mask=mask, noise=noise, bias=bias, causal=causal, left_only=causal,
                                    mask_queries=causal)
    assert rel_l2(out.cpu(), ref.cpu()) < 2e-5


@pytest.mark.parametrize('seq_shape,window,ext,chunk,causal,with_mask,d,dtype', [
    ((14, 14), 7, 0, 2, False, False, 64, torch.float32), ((28, 28), 7, 0, 4, False, False, 64, torch.float32),
    ((14, 14), 7, 3, 2, False, False, 64, torch.float32), ((96,), 16, 8, 12, False, True, 64, torch.float32),
    ((128,), 32, 0, 16, True, True, 64, torch.float32), ((96,), 16, 16, 8, True, False, 64, torch.float32),
    ((14, 14), 7, 0, 2, False, False, 32, torch.float32), ((96,), 16, 8, 12, False, True, 128, torch.float32),
    ((28, 28), 7, 0, 4, False, False, 16, torch.float32), ((512,), 128, 128, 64, True, True, 64, torch.float32),
    ((28, 28), 7, 0, 4, False, False, 64, torch.float16), ((512,), 128, 128, 64, True, False, 64, torch.bfloat16)])
def test_backward_kernels_equal_autograd_through_the_recomputation(seq_shape, window, ext, chunk, causal, with_mask, d, dtype):
    """`eva_backward` (csrc/eva_backward.cu) against autograd through `eva_core_torch` on identical inputs: gradients with respect
    to q, k, v, the bias table and all eight adaptive Linear / LayerNorm parameters."""
    from efficient_attention import _recompute
    from test_gpu_parity import _rand_ada
    dev = _dev()
    B, H = 2, 2
    N = math.prod(seq_shape)
    g = torch.Generator().manual_seed(N + window + ext + d)
    qkv = torch.randn(B, N, 3, H, d, generator=g).to(dev, dtype)
    two_d = len(seq_shape) == 2
    L = window * window if two_d else window
    J = (window + 2 * ext) ** 2 if two_d else window + (ext if causal else 2 * ext)
    bias = (0.5 * torch.randn(H, L, J, generator=g)).to(dev)
    ada = {k_: (v_.to(dev) if v_ is not None else None) for k_, v_ in _rand_ada(d, g).items()}
    mask = None
    if with_mask:
        mask = torch.zeros(B, N, dtype=torch.bool, device=dev)
        mask[1, N - 9:] = True
    chunk_ext = 0 if causal else ext
    noise = torch.randn(B, H, _recompute.num_chunks_of(seq_shape, chunk), d, generator=g).to(dev)
    geometry = dict(seq_shape=seq_shape, window=window, ext=ext, chunk=chunk, chunk_ext=chunk_ext, causal=causal,
                    halo_left_only=causal, mask_queries=causal)
    names = ('wq', 'bq', 'gq', 'betq', 'wk', 'bk', 'gk', 'betk')
    w = torch.randn(B, N, H * d, generator=g).to(dev)

    def run(impl):
        prev = _recompute.set_backward_impl(impl)
        try:
            x = qkv.clone().requires_grad_(True)
            b_ = bias.clone().requires_grad_(True)
            prm = [ada[n_].clone().requires_grad_(True) for n_ in names]
            out = _recompute.eva_core(x[:, :, 0], x[:, :, 1], x[:, :, 2], geometry=geometry, mu_coeff=1.0 if causal else 0.5,
                                      params=prm, pad_mask=mask, noise=noise, bias=b_)
            (out.float() * w).sum().backward()
            return [x.grad.float(), b_.grad] + [p_.grad for p_ in prm]
        finally:
            _recompute.set_backward_impl(prev)
    got, want = run('cuda'), run('torch')
    tol = 2e-5 if dtype == torch.float32 else 6e-3     # 16-bit: the saved forward output and grad_out are rounded to the I/O format
    for n_, a_, b_ in zip(('qkv', 'bias') + names, got, want):
        assert torch.isfinite(a_).all(), n_
        assert rel_l2(a_.cpu(), b_.cpu()) < tol, (n_, rel_l2(a_.cpu(), b_.cpu()))


@pytest.mark.parametrize('seq_shape,window,chunk,dtype,with_bias,B', [
    ((28, 28), 7, 4, torch.float16, True, 2), ((14, 14), 7, 2, torch.bfloat16, True, 3), ((28, 28), 7, 4, torch.float16, False, 40),
    ((96,), 16, 8, torch.float16, True, 2), ((56, 56), 7, 8, torch.bfloat16, False, 2), ((16, 16), 8, 2, torch.float16, True, 2)])
def test_tcgen05_backward_kernel_matches_the_cuda_core_kernel(seq_shape, window, chunk, dtype, with_bias, B):
    """eva_bwd_sm100.cu (tensor cores, 16-bit P / dS tiles) against eva_backward.cu (float32) on identical 16-bit inputs, and the
    dispatch counter proves which one ran."""
    from efficient_attention import _abi, _recompute
    from test_gpu_parity import _rand_ada
    dev = _dev()
    lib = _abi.load()
    H, d = 3, 64
    N = math.prod(seq_shape)
    g = torch.Generator().manual_seed(N + window + chunk)
    qkv = torch.randn(B, N, 3, H, d, generator=g).to(dev, dtype)
    L = window * window if len(seq_shape) == 2 else window
    bias = (0.5 * torch.randn(H, L, L, generator=g)).to(dev) if with_bias else None
    ada = {k_: v_.to(dev) for k_, v_ in _rand_ada(d, g).items()}
    noise = torch.randn(B, H, _recompute.num_chunks_of(seq_shape, chunk), d, generator=g).to(dev)
    geometry = dict(seq_shape=seq_shape, window=window, ext=0, chunk=chunk, chunk_ext=0)
    names = ('wq', 'bq', 'gq', 'betq', 'wk', 'bk', 'gk', 'betk')
    w = torch.randn(B, N, H * d, generator=g).to(dev)

    def run(mode):
        lib.eva_debug_set_bwd_tc(mode)
        try:
            before = lib.eva_debug_bwd_tc_count()
            x = qkv.clone().requires_grad_(True)
            b_ = bias.clone().requires_grad_(True) if with_bias else None
            prm = [ada[n_].clone().requires_grad_(True) for n_ in names]
            out = _recompute.eva_core(x[:, :, 0], x[:, :, 1], x[:, :, 2], geometry=geometry, mu_coeff=0.5, params=prm, noise=noise, bias=b_,
                                      packed=x if mode else None)       # mode 1 also takes the packed-gradient route (grad_qkv_io)
            (out.float() * w).sum().backward()
            assert lib.eva_debug_bwd_tc_count() - before == mode
            return [x.grad.float()] + ([b_.grad] if with_bias else []) + [p_.grad for p_ in prm]
        finally:
            lib.eva_debug_set_bwd_tc(-1)
    got, want = run(1), run(0)
    for n_, a_, b_ in zip(('qkv',) + (('bias',) if with_bias else ()) + names, got, want):
        assert torch.isfinite(a_).all(), n_
        assert rel_l2(a_.cpu(), b_.cpu()) < (3e-3 if dtype == torch.float16 else 1.2e-2), (n_, rel_l2(a_.cpu(), b_.cpu()))


@pytest.mark.parametrize('seq_shape,window,chunk,dtype,expect_path', [
    ((28, 28), 7, 4, torch.float16, 1), ((14, 14), 7, 2, torch.bfloat16, 1), ((96,), 16, 8, torch.float32, 0)])
def test_forward_keeps_the_chunk_statistics_for_the_backward(seq_shape, window, chunk, dtype, expect_path):
    """`EvaGeometry.keep_stats`: k_bar | beta left in the workspace by whichever forward kernel ran equal `eva_chunk_stats`."""
    from efficient_attention import _abi, _recompute
    from test_gpu_parity import _rand_ada
    dev = _dev()
    B, H, d = 3, 3, 64
    N = math.prod(seq_shape)
    g = torch.Generator().manual_seed(N)
    qkv = torch.randn(B, N, 3, H, d, generator=g).to(dev, dtype)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    ada_t = {k_: v_.to(dev) for k_, v_ in _rand_ada(d, g).items()}
    ada = _abi.adaptive(*[ada_t[n] for n in ('wq', 'bq', 'gq', 'betq', 'wk', 'bk', 'gk', 'betk')], mu_coeff=0.5)
    noise = torch.randn(B, H, _recompute.num_chunks_of(seq_shape, chunk), d, generator=g).to(dev)
    geometry = dict(seq_shape=seq_shape, window=window, ext=0, chunk=chunk, chunk_ext=0)
    out, path, stats = _abi.eva_forward(q, k, v, _abi.eva_geometry(q, keep_stats=True, **geometry), ada, noise=noise,
                                        return_path=True, return_stats=True)
    k_bar, beta = stats[:2]
    assert path == expect_path
    ref_out = _abi.eva_forward(q, k, v, _abi.eva_geometry(q, **geometry), ada, noise=noise)
    assert torch.equal(out, ref_out)
    kb, bt = _abi.eva_chunk_stats(q, k, v, _abi.eva_geometry(q, **geometry), ada, noise=noise)
    tol = 1e-6 if dtype == torch.float32 else 4e-3
    assert rel_l2(k_bar.cpu(), kb.cpu()) < tol and rel_l2(beta.cpu(), bt.cpu()) < tol, (rel_l2(k_bar.cpu(), kb.cpu()), rel_l2(beta.cpu(), bt.cpu()))


@pytest.mark.parametrize('seq_shape,window,ext,chunk,causal,with_mask,dtype', [
    ((14, 14), 7, 3, 2, False, False, torch.float16), ((96,), 16, 8, 12, False, True, torch.float16),
    ((512,), 128, 128, 64, True, True, torch.float16), ((512,), 256, 0, 256, True, False, torch.bfloat16),
    ((14, 14), 7, 0, 0, False, False, torch.float16), ((197,), 197, 0, 0, False, True, torch.float16),
    ((400,), 400, 0, 0, False, False, torch.bfloat16), ((28, 28), 14, 0, 4, False, False, torch.float16)])
def test_generic_tcgen05_backward_kernel_matches_the_cuda_core_kernel(seq_shape, window, ext, chunk, causal, with_mask, dtype):
    """eva_window_bwd_gen_sm100.cu (any geometry, head_dim 64, 16-bit) against window_attn_bwd_kernel (float32 CUDA cores) on
    identical 16-bit inputs: gradients of q, k, v, the bias table and (with chunks) the adaptive parameters."""
    from efficient_attention import _abi, _recompute
    from test_gpu_parity import _rand_ada
    dev = _dev()
    lib = _abi.load()
    B, H, d = 2, 2, 64
    N = math.prod(seq_shape)
    g = torch.Generator().manual_seed(N + window + ext + chunk)
    qkv = torch.randn(B, N, 3, H, d, generator=g).to(dev, dtype)
    two_d = len(seq_shape) == 2
    L = window * window if two_d else window
    J = (window + 2 * ext) ** 2 if two_d else window + (ext if causal else 2 * ext)
    bias = (0.5 * torch.randn(H, L, J, generator=g)).to(dev) if L * J <= 1 << 16 else None
    mask = None
    if with_mask:
        mask = torch.zeros(B, N, dtype=torch.bool, device=dev)
        mask[1, N - 9:] = True
    names = ('wq', 'bq', 'gq', 'betq', 'wk', 'bk', 'gk', 'betk')
    ada = {k_: v_.to(dev) for k_, v_ in _rand_ada(d, g).items()}
    w = torch.randn(B, N, H * d, generator=g).to(dev)
    if mask is not None and causal:
        w = w * (~mask).unsqueeze(-1)                     # padded queries of the causal layer are don't-care rows
    softmax_like = chunk == 0 and window == N
    if chunk:
        noise = torch.randn(B, H, _recompute.num_chunks_of(seq_shape, chunk), d, generator=g).to(dev)
        geometry = dict(seq_shape=seq_shape, window=window, ext=ext, chunk=chunk, chunk_ext=0 if causal else ext, causal=causal,
                        halo_left_only=causal, mask_queries=causal)
    else:
        geometry = dict(seq_shape=seq_shape, window=window, ext=ext, chunk=0, chunk_ext=0, mask_is_neg_inf=softmax_like)

    def run(mode):
        lib.eva_debug_set_bwd_tc(mode)
        try:
            before = lib.eva_debug_bwd_tc_count()
            x = qkv.clone().requires_grad_(True)
            b_ = bias.clone().requires_grad_(True) if bias is not None else None
            prm = [ada[n_].clone().requires_grad_(True) for n_ in names]
            if chunk:
                out = _recompute.eva_core(x[:, :, 0], x[:, :, 1], x[:, :, 2], geometry=geometry, mu_coeff=1.0 if causal else 0.5,
                                          params=prm, pad_mask=mask, noise=noise, bias=b_)
            else:
                out = _recompute.window_core(x[:, :, 0], x[:, :, 1], x[:, :, 2], geometry=geometry, pad_mask=mask, bias=b_)
            (out.float() * w).sum().backward()
            assert lib.eva_debug_bwd_tc_count() - before == mode
            return [x.grad.float()] + ([b_.grad] if b_ is not None else []) + ([p_.grad for p_ in prm] if chunk else [])
        finally:
            lib.eva_debug_set_bwd_tc(-1)
    got, want = run(1), run(0)
    labels = ('qkv',) + (('bias',) if bias is not None else ()) + (names if chunk else ())
    for n_, a_, b_ in zip(labels, got, want):
        assert torch.isfinite(a_).all(), n_
        assert rel_l2(a_.cpu(), b_.cpu()) < (3e-3 if dtype == torch.float16 else 1.5e-2), (n_, rel_l2(a_.cpu(), b_.cpu()))


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-5), (torch.float16, 2e-2), (torch.bfloat16, 6e-2)])
def test_lara_fused_backward_steps_match_the_explicit_formulas(dtype, tol):
    """`_lara_stage2_backward_fused` (three `lara_backward_step` kernels between batched GEMMs in the activation format) against
    `_lara_stage2_backward` (the same formulas in float32 PyTorch ops) on random inputs."""
    from efficient_attention import _recompute
    dev = _dev()
    B, H, N, C, d = 2, 3, 196, 49, 64
    g = torch.Generator().manual_seed(7)
    q, k, v, go = (torch.randn(B, H, N, d, generator=g).to(dev) for _ in range(4))
    q, k, v, go = (t.to(dtype).float() for t in (q, k, v, go))                     # identical values on both sides
    q_bar, omega = (0.5 * torch.randn(B, H, C, d, generator=g).to(dev) for _ in range(2))
    lp = torch.randn(B, H, C, 1, generator=g).to(dev)
    bh = torch.softmax(torch.randn(B, H, C, 1, generator=g), 2).to(dev)
    want = _recompute._lara_stage2_backward(q, k, v, go, q_bar, omega, lp, bh, 2.0)
    flat = lambda t: t.reshape(B * H, *t.shape[2:]).contiguous()
    got = _recompute._lara_stage2_backward_fused(flat(q).to(dtype), flat(k).to(dtype), flat(v).to(dtype), flat(go).to(dtype), flat(q_bar), flat(omega),
                                                 flat(lp).squeeze(-1), flat(bh).squeeze(-1), 2.0)
    for name, a_, b_ in zip(('dq', 'dk', 'dv', 'dq_bar', 'domega', 'dlp', 'dbh'), got, want):
        assert torch.isfinite(a_).all(), name
        assert rel_l2(a_.reshape(-1).cpu(), b_.reshape(-1).cpu()) < tol, (name, rel_l2(a_.reshape(-1).cpu(), b_.reshape(-1).cpu()))


def _grads_of(module, cfg, a, dev, dtype):
    """loss = <y, w> for a fixed w; returns y, dL/dx and {name: dL/dparam}."""
    x = a['x'].to(device=dev, dtype=dtype).requires_grad_(True)
    mask = a['mask'].to(dev) if a['mask'] is not None else None
    noise = a['noise'].to(device=dev, dtype=torch.float32) if a['noise'] is not None else None
    if cfg['kind'] == 'causal_eva':
        y = module(x, x, x, key_padding_mask=mask, noise=noise)[0]
    elif cfg['kind'] in ('local', 'softmax'):
        y = module(x, mask)
    else:
        y = module(x, mask, noise=noise)
    w = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).to(device=dev, dtype=y.dtype)
    (y * w).sum().backward()
    return y.detach(), x.grad.detach(), {n: p.grad.detach() for n, p in module.named_parameters() if p.grad is not None}


def _oracle_grads(cfg, sd, a):
    sd64 = {k_: (v_.double().requires_grad_(True) if v_.is_floating_point() else v_) for k_, v_ in sd.items()}
    x = a['x'].double().requires_grad_(True)
    noise = a['noise'].double() if a['noise'] is not None else None
    if cfg['kind'] == 'causal_eva':
        y = O.causal_eva_forward(sd64, cfg, x, pad_mask=a['mask'], noise=noise)
    elif cfg['kind'] == 'lara':
        y = O.lara_forward(sd64, cfg, x, pad_mask=a['mask'], noise=noise)
    elif cfg['kind'] == 'local':
        y = O.local_forward(sd64, cfg, x, pad_mask=a['mask'])
    elif cfg['kind'] == 'softmax':
        y = O.softmax_forward(sd64, cfg, x, pad_mask=a['mask'])
    else:
        y = O.eva_forward(sd64, cfg, x, pad_mask=a['mask'], noise=noise)
    w = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).double()
    (y * w).sum().backward()
    return y.detach(), x.grad, {k_: v_.grad for k_, v_ in sd64.items() if v_.is_floating_point() and v_.grad is not None}


@pytest.mark.parametrize('name', ['eva_2d_train', 'eva_1d_train', 'eva_c1_train', 'eva_c3_train', 'causal_numchunks_train',
                                  'lara_c4_train', 'lara_2d_train_anti', 'lara_2d_train_multi', 'lara_1d_even', 'lara_1d_uneven_mask',
                                  'lara_2d_dense', 'lara_2d_dense_vmixed', 'lara_2d_vmixed_biased', 'local_1d_mask', 'local_2d_rpe',
                                  'softmax_mask'])
def test_module_gradients_match_oracle_autograd_fp32(name):
    cfg, sd, a = load_golden(name, dtype=torch.float32)
    m = build_module(cfg)
    m.load_state_dict(sd)
    m = m.to(_dev())
    m.train(a['noise'] is not None)              # fixtures without a noise draw were generated in eval mode
    y, gx, gp = _grads_of(m, cfg, a, _dev(), torch.float32)
    y64, gx64, gp64 = _oracle_grads(cfg, sd, a)
    assert rel_l2(y.cpu(), y64) < 2e-5
    assert rel_l2(gx.cpu(), gx64) < 1e-4, ('x', rel_l2(gx.cpu(), gx64))
    assert set(gp) == set(gp64), (sorted(set(gp) ^ set(gp64)))
    for n_ in gp64:
        if float(gp64[n_].norm()) > 1e-9:
            assert rel_l2(gp[n_].cpu(), gp64[n_]) < 2e-4, (n_, rel_l2(gp[n_].cpu(), gp64[n_]))


def test_autocast_training_step_through_the_fused_path():
    """DeiT-style step (vit/engine.py:47-62): fp16 autocast forward through the fused tcgen05 kernel, backward by recomputation,
    every parameter receives a finite gradient; the forward took path 1 or 3."""
    import bench
    from test_gpu_parity import _path_counts
    m = bench.build_layer(_dev(), torch.float32).train()
    x = torch.randn(8, 28, 28, 192, device=_dev(), requires_grad=True)
    before = _path_counts()
    with torch.autocast('cuda', dtype=torch.float16):
        y = m(x)
    after = _path_counts()
    assert after[0] == before[0] and (after[1] - before[1]) + (after[3] - before[3]) == 1
    y.float().pow(2).mean().backward()
    assert torch.isfinite(x.grad).all() and float(x.grad.abs().sum()) > 0
    for n_, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n_


def test_causal_attention_dropout():
    """causal_eva.py:778: dropout on the joint probabilities.  With an all-keep mask the output before the projection bias is the
    no-dropout output / (1 - p); with random draws it is unbiased; gradients flow."""
    from argparse import Namespace
    import efficient_attention as ea
    torch.manual_seed(0)
    ns = Namespace(adaptive_proj='qk', num_chunks=None, chunk_size=16, causal=True, use_t5_rpe=True, window_size=32, overlap_window=False)
    m = ea.CausalEVAttention(128, 2, dropout=0.25, self_attention=True, attn_args=ns).to(_dev())
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 2:
                p.normal_(0, 1.0 / math.sqrt(p.shape[1]))
    x = torch.randn(64, 2, 128, device=_dev())
    noise = torch.randn(2, 2, 4, 64, device=_dev())
    m.eval()
    with torch.no_grad():
        y0 = m(x, x, x, noise=noise)[0]
    m.train()
    keep = torch.ones(2, 2, 2, 32, 32 + 4, dtype=torch.bool, device=_dev())
    with torch.no_grad():
        y1 = m(x, x, x, noise=noise, drop_mask=keep)[0]
    b = m.out_proj.bias
    assert rel_l2(((y1 - b) * 0.75).cpu(), (y0 - b).cpu()) < 2e-5
    with torch.no_grad():
        ys = torch.stack([m(x, x, x, noise=noise)[0] for _ in range(300)]).mean(0)
    assert rel_l2(ys.cpu(), y0.cpu()) < 0.08                   # unbiased: the mean over draws approaches the no-dropout output
    xg = x.clone().requires_grad_(True)
    m(xg, xg, xg, noise=noise)[0].pow(2).mean().backward()
    assert torch.isfinite(xg.grad).all() and float(xg.grad.abs().sum()) > 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_ddp_two_rank_training_step():
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                          '--master-port', '29641', os.path.join(ROOT, 'tools', 'ddp_smoke.py')], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert 'DDP_SMOKE_OK' in out.stdout, out.stdout[-2000:]
