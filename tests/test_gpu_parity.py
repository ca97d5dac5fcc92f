"""GPU parity tests (`-m gpu`): the CUDA path, called through the C ABI (ctypes) and through the
drop-in modules, against the golden fixtures and the CPU oracle on identical inputs.

Tolerances (relative L2 per output tensor, reference evaluated in float64):
  float32 I/O                      <= 2e-5   (fp32 math on both sides of the softmax)
  float16 I/O, core only           <= 1e-3   (north_star tolerance; output rounding alone is ~2e-4)
  bfloat16 I/O, core only          <= 6e-3   (bf16 output store alone costs ~1.7e-3; reported, not the target)
"""
import math

import pytest
import torch

from conftest import golden_names, load_golden, rel_l2
from helpers import build_module, run_module
from oracle import eva_oracle as O

pytestmark = pytest.mark.gpu

TOL_F32 = 2e-5
TOL_F16 = 1e-3
TOL_BF16 = 6e-3


def _dev():
    return torch.device('cuda', 0)


def _rand_ada(d, g, ln=True):
    wq, wk = torch.randn(d, d, generator=g) / math.sqrt(d), torch.randn(d, d, generator=g) / math.sqrt(d)
    bq, bk = 0.1 * torch.randn(d, generator=g), 0.1 * torch.randn(d, generator=g)
    if not ln:
        return dict(wq=wq, bq=bq, gq=None, betq=None, wk=wk, bk=bk, gk=None, betk=None)
    return dict(wq=wq, bq=bq, gq=1 + 0.1 * torch.randn(d, generator=g), betq=0.1 * torch.randn(d, generator=g),
                wk=wk, bk=bk, gk=1 + 0.1 * torch.randn(d, generator=g), betk=0.1 * torch.randn(d, generator=g))


def _abi_ada(p, dev, mu_coeff):
    from efficient_attention import _abi
    mv = lambda t: None if t is None else t.to(dev)
    return _abi.adaptive(mv(p['wq']), mv(p['bq']), mv(p['gq']), mv(p['betq']), mv(p['wk']), mv(p['bk']),
                         mv(p['gk']), mv(p['betk']), mu_coeff=mu_coeff)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', golden_names())
def test_module_matches_reference_golden_fp32(name):
    cfg, sd, a = load_golden(name, dtype=torch.float32)
    m = build_module(cfg)
    m.load_state_dict(sd)
    m = m.to(_dev())
    m.train(a['noise'] is not None)
    y = run_module(m, cfg, a, _dev(), torch.float32)
    assert y.shape == a['y'].shape
    err = rel_l2(y.cpu(), a['y'])
    assert err < TOL_F32, (name, err)


# ---------------------------------------------------------------------------------------------
CORE_CASES = {
    # name: (B, H, seq_shape, d, window, ext, chunk)
    'c1_14x14': (2, 3, (14, 14), 64, 7, 0, 2),
    'c3_28x28': (2, 3, (28, 28), 64, 7, 0, 4),
    'pvt_d32_56x56': (1, 2, (56, 56), 32, 7, 0, 8),
    'overlap_14x14': (1, 2, (14, 14), 64, 7, 3, 2),
    'seq_1d_d128': (2, 2, (96,), 128, 16, 8, 12),
}


@pytest.mark.parametrize('dtype,tol', [(torch.float32, TOL_F32), (torch.float16, TOL_F16), (torch.bfloat16, TOL_BF16)])
@pytest.mark.parametrize('case', sorted(CORE_CASES))
def test_eva_core_through_cabi_vs_oracle(case, dtype, tol):
    from efficient_attention import _abi
    B, H, shape, d, w, ext, chunk = CORE_CASES[case]
    N = math.prod(shape)
    g = torch.Generator().manual_seed(hash(case) % 1000)
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.2).to(dtype)        # quantise once; both sides see the same values
    two_d = len(shape) == 2
    L = w * w if two_d else w
    J = (w + 2 * ext) ** 2 if two_d else w + 2 * ext
    bias = 0.5 * torch.randn(H, L, J, generator=g)
    ada = _rand_ada(d, g)
    mask = None
    if not two_d:
        mask = torch.zeros(B, N, dtype=torch.bool)
        mask[1, N - 10:] = True
    q64, k64, v64 = (qkv[:, :, i].permute(0, 2, 1, 3).double() for i in range(3))
    want, kbar_w, beta_w = O.eva_core(q64, k64, v64, seq_shape=shape, window=w, ext=ext, chunk=chunk, chunk_ext=ext,
                                      **{k_: (v_.double() if v_ is not None else None) for k_, v_ in ada.items()},
                                      mu_coeff=0.5, pad_mask=mask, bias=bias.double(), return_stats=True)
    dev = _dev()
    qkv_d = qkv.to(dev)
    q, k, v = qkv_d[:, :, 0], qkv_d[:, :, 1], qkv_d[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=shape, window=w, ext=ext, chunk=chunk, chunk_ext=ext)
    ada_s = _abi_ada(ada, dev, 0.5)
    mask_d = mask.to(dev) if mask is not None else None
    # stage A alone
    kbar, beta = _abi.eva_chunk_stats(q, k, v, geom, ada_s, pad_mask=mask_d)
    assert rel_l2(kbar.cpu(), kbar_w) < 2e-5 and rel_l2(beta.cpu(), beta_w) < 2e-5
    # stage B alone, fed with the oracle's statistics
    out_b = _abi.eva_window_attention(q, k, v, geom, k_bar=kbar_w.float().to(dev), beta=beta_w.float().to(dev),
                                      pad_mask=mask_d, bias=bias.to(dev))
    # both stages in one call
    out, path = _abi.eva_forward(q, k, v, geom, ada_s, pad_mask=mask_d, bias=bias.to(dev), return_path=True)
    want_flat = want.permute(0, 2, 1, 3).reshape(B, N, H * d)
    for name, o in (('stage_b', out_b), ('forward', out)):
        err = rel_l2(o.cpu(), want_flat)
        assert err < tol, (case, dtype, name, path, err)


def test_c5_causal_layer_full_shape_fp32():
    """BASELINE config c5 at full size: T=4096, B=2, C=512, h=8, window=chunk=256, causal, T5 bias."""
    from argparse import Namespace
    import efficient_attention as ea
    torch.manual_seed(3)
    m = ea.CausalEVAttention(512, 8, self_attention=True, attn_args=Namespace(
        adaptive_proj='qk', num_chunks=None, chunk_size=256, causal=True, use_t5_rpe=True, window_size=256,
        overlap_window=False)).eval()
    with torch.no_grad():
        m.rel_pos_bias.relative_attention_bias.weight.normal_(0, 0.5)
    x = torch.randn(4096, 2, 512)
    cfg = dict(num_heads=8, window_size=256, overlap_window=False, chunk_size=256, num_chunks=None, causal=True,
               use_t5_rpe=True, adaptive_proj='qk')
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    want = O.causal_eva_forward(sd, cfg, x.double())
    with torch.no_grad():
        got = m.to(_dev())(x.to(_dev()), None, None)[0]
    assert rel_l2(got.cpu(), want) < TOL_F32


def test_c4_lara_layer_full_shape_fp32():
    """BASELINE config c4 at full width: C=384, h=6, 14x14, 49 landmarks, pool-mixed, mis-opt."""
    import efficient_attention as ea
    torch.manual_seed(4)
    m = ea.AttentionFactory.build_attention('lara', dict(dim=384, num_heads=6, num_landmarks=49,
                                                         proposal_gen='pool-mixed', mis_type='mis-opt')).eval()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 2:
                p.normal_(0, 1.0 / math.sqrt(p.shape[1]))
    x = torch.randn(4, 14, 14, 384)
    cfg = dict(num_heads=6, num_landmarks=49, proposal_gen='pool-mixed', mis_type='mis-opt', alpha_coeff=1.0)
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    want = O.lara_forward(sd, cfg, x.double())
    with torch.no_grad():
        got = m.to(_dev())(x.to(_dev()))
    assert rel_l2(got.cpu(), want) < TOL_F32


# ---- size-independent properties at the benchmark's shape --------------------------------------
def _bench_layer(dtype):
    import bench
    return bench.build_layer(_dev(), dtype)


def test_batch_permutation_equivariance_and_determinism_fp16():
    m = _bench_layer(torch.float16)
    torch.manual_seed(5)
    x = torch.randn(64, 28, 28, 192, device=_dev(), dtype=torch.float16)
    perm = torch.randperm(64, device=_dev())
    with torch.no_grad():
        y1, y2, yp = m(x), m(x), m(x[perm])
    assert torch.equal(y1, y2)                     # run-to-run bitwise deterministic
    assert torch.equal(y1[perm], yp)               # every (batch, head) is an independent unit


def test_head_independence_fp32():
    """Heads only meet in qkv / proj; with identity proj, perturbing head 0's input slice of the qkv
    output must leave the other heads' core outputs bit-identical."""
    from efficient_attention import _abi
    torch.manual_seed(6)
    dev = _dev()
    g = torch.Generator().manual_seed(6)
    qkv = torch.randn(2, 196, 3, 3, 64, generator=g).to(dev)
    ada = _abi_ada(_rand_ada(64, g), dev, 0.5)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(14, 14), window=7, ext=0, chunk=2, chunk_ext=0)
    o1 = _abi.eva_forward(q, k, v, geom, ada).view(2, 196, 3, 64)
    qkv2 = qkv.clone()
    qkv2[:, :, :, 0] += 1.0
    o2 = _abi.eva_forward(qkv2[:, :, 0], qkv2[:, :, 1], qkv2[:, :, 2], geom, ada).view(2, 196, 3, 64)
    assert torch.equal(o1[:, :, 1:], o2[:, :, 1:]) and not torch.equal(o1[:, :, 0], o2[:, :, 0])


def test_padding_invariance_1d():
    """Appending masked tokens that fill whole extra windows must not change the valid outputs' local
    part; with chunk keys present the chunk partition changes, so compare pure local attention."""
    cfg, sd, a = load_golden('local_1d_mask', dtype=torch.float32)
    m = build_module(cfg)
    m.load_state_dict(sd)
    m = m.to(_dev()).eval()
    x = a['x'].to(_dev())
    with torch.no_grad():
        y = m(x, None)
        x_long = torch.cat([x, torch.randn(2, 16, x.shape[-1], device=_dev())], 1)
        mask = torch.zeros(2, x_long.shape[1], dtype=torch.bool, device=_dev())
        mask[:, x.shape[1]:] = True
        y_long = m(x_long, mask)
    # tokens 0..23 live in windows whose halo never reaches the appended region (window 8, halo 4)
    assert rel_l2(y_long[:, :24].cpu(), y[:, :24].cpu()) < 1e-6


def test_causal_consistency_on_gpu():
    """causal_eva.py:916-950 on the device: a prefix's outputs equal the full sequence's."""
    cfg, sd, a = load_golden('causal_selfcheck', dtype=torch.float32)
    m = build_module(cfg)
    m.load_state_dict(sd)
    m = m.to(_dev()).eval()
    x = a['x'].to(_dev())
    with torch.no_grad():
        full = m(x, None, None)[0]
        for t in (26, 64, 65, 100):
            part = m(x[:t], None, None)[0]
            assert torch.allclose(part[25], full[25], atol=2e-5, rtol=0), t


def test_cabi_error_paths_on_device():
    from efficient_attention import _abi
    dev = _dev()
    qkv = torch.randn(1, 196, 3, 2, 64, device=dev)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(14, 14), window=7, ext=0, chunk=2, chunk_ext=0)
    g = torch.Generator().manual_seed(0)
    ada = _abi_ada(_rand_ada(64, g), dev, 0.5)
    with pytest.raises(_abi.EvaKernelError, match='bias_stride_h'):
        _abi.eva_forward(q, k, v, geom, ada, bias=torch.zeros(2, 49, 48, device=dev))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _abi.eva_forward(q.cpu(), k.cpu(), v.cpu(), geom, ada)
    misaligned = torch.randn(1, 196, 3, 2, 66, device=dev)[..., 1:65]
    with pytest.raises(AssertionError):
        _abi.heads_view(misaligned[:, :, 0].transpose(2, 3))


def test_host_pipeline_matches_direct_forward():
    """efficient_attention.streaming.HostPipeline (chunked H2D / forward / D2H on three streams) must give the
    same bits as one forward over the whole batch."""
    from efficient_attention.streaming import HostPipeline
    m = _bench_layer(torch.float16)
    torch.manual_seed(7)
    x_host = torch.randn(50, 28, 28, 192).half().pin_memory()
    y_host = torch.empty_like(x_host).pin_memory()
    pipe = HostPipeline(m, chunk=16)
    for _ in range(2):            # second pass re-uses the staging buffers
        pipe(x_host, y_host)
    torch.cuda.synchronize()
    with torch.no_grad():
        want = m(x_host.to(_dev())).cpu()
    assert torch.equal(y_host, want)
    # deferred join: two batches in flight over two host buffer pairs, host-side wait on pipe.done
    pipe2 = HostPipeline(m, chunk=16, defer_join=True)
    xs = [x_host, (x_host * 0.5).pin_memory()]
    ys = [torch.empty_like(x_host).pin_memory() for _ in range(2)]
    events = []
    for i in range(4):
        pipe2(xs[i & 1], ys[i & 1])
        events.append(pipe2.done)
    for e in events:
        e.synchronize()
    with torch.no_grad():
        want2 = m(xs[1].to(_dev())).cpu()
    assert torch.equal(ys[0], want) and torch.equal(ys[1], want2)


def test_host_pipeline_first_call_after_other_work():
    """A pipeline's FIRST call allocates its staging buffers while the compute stream is busy: blocks recycled by the caching
    allocator must not be overwritten by the copy-in stream before the kernels that still use them have run (regression: the
    H2D copy of chunk 1 used to land in a block the forward of chunk 0 had just released)."""
    from efficient_attention.streaming import HostPipeline
    m = _bench_layer(torch.float16)
    torch.manual_seed(9)
    B = 384
    x_host = torch.randn(B, 28, 28, 192).half().pin_memory()
    y_host = torch.empty_like(x_host).pin_memory()
    with torch.no_grad():
        want = m(x_host.to(_dev())).cpu()
    for trial in range(4):
        with torch.no_grad():
            for _ in range(2):
                m(x_host[:96 * (trial + 1)].to(_dev()))      # leaves freed blocks of various sizes behind, kernels still queued
        HostPipeline(m, chunk=48)(x_host, y_host)              # dropped right after the call
        torch.cuda.synchronize()
        assert torch.equal(y_host, want), trial


class _cluster_mode:
    """Run a block with the opt-in cluster-resident kernel (eva_cluster_sm100.cu) switched on / off."""

    def __init__(self, on):
        self.on = on

    def __enter__(self):
        import ctypes
        from efficient_attention import _abi
        self.lib = _abi.load()
        self.lib.eva_debug_set_cluster_mode.restype = ctypes.c_int
        self.prev = self.lib.eva_debug_set_cluster_mode(ctypes.c_int(1 if self.on else 0))

    def __exit__(self, *exc):
        import ctypes
        self.lib.eva_debug_set_cluster_mode(ctypes.c_int(self.prev))


@pytest.mark.parametrize('grid,chunk,cluster', [(14, 2, False), (28, 4, False), (28, 4, True)])
@pytest.mark.parametrize('adaptive', ['default', 'no-ln', 'none'])
@pytest.mark.parametrize('with_noise,with_bias', [(False, True), (True, False), (True, True)])
def test_fused_kernel_variants_fp16(grid, chunk, cluster, adaptive, with_noise, with_bias):
    """Every option the fused tcgen05/TMA kernel implements (adaptive_proj variants, training noise, optional
    bias) on both instantiated geometries, against the oracle on identical fp16 inputs.  The call must take
    the fused path (path == 1): a silent fall-back to the generic kernels would hide a regression."""
    from efficient_attention import _abi
    B, H, d, w = 3, 3, 64, 7
    N = grid * grid
    g = torch.Generator().manual_seed(grid * 100 + len(adaptive) * 10 + int(with_noise) * 2 + int(with_bias))
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.1).half()
    bias = 0.5 * torch.randn(H, 49, 49, generator=g) if with_bias else None
    ada = _rand_ada(d, g, ln=(adaptive != 'no-ln'))
    if adaptive == 'none':
        ada.update(wq=None, bq=None, gq=None, betq=None)
    noise = torch.randn(B, H, 49, d, generator=g) if with_noise else None
    q64, k64, v64 = (qkv[:, :, i].permute(0, 2, 1, 3).double() for i in range(3))
    want = O.eva_core(q64, k64, v64, seq_shape=(grid, grid), window=w, ext=0, chunk=chunk, chunk_ext=0,
                      **{k_: (v_.double() if v_ is not None else None) for k_, v_ in ada.items()}, mu_coeff=0.5,
                      use_q=(adaptive != 'none'), noise=noise.double() if with_noise else None,
                      bias=bias.double() if with_bias else None)
    want = want.permute(0, 2, 1, 3).reshape(B, N, H * d)
    dev = _dev()
    qd = qkv.to(dev)
    geom = _abi.eva_geometry(qd[:, :, 0], seq_shape=(grid, grid), window=w, ext=0, chunk=chunk, chunk_ext=0)
    with _cluster_mode(cluster):
        out, path = _abi.eva_forward(qd[:, :, 0], qd[:, :, 1], qd[:, :, 2], geom, _abi_ada(ada, dev, 0.5),
                                     noise=noise.to(dev) if with_noise else None, bias=bias.to(dev) if with_bias else None,
                                     return_path=True)
    assert path == (3 if cluster else 1)          # 3: cluster-resident kernel (opt-in), 1: streamed kernel
    err = rel_l2(out.cpu(), want)
    assert err < TOL_F16, (grid, adaptive, with_noise, with_bias, err)


@pytest.mark.parametrize('grid,chunk,B,cluster', [(14, 2, 400, False), (28, 4, 120, False), (28, 4, 200, True)])
def test_fused_kernel_many_items_per_cta_fp16(grid, chunk, B, cluster):
    """More (batch, head) items than resident CTAs (2 x 148): every CTA loops over several items handed out by the
    dynamic work counter, which exercises the ring / barrier phase bookkeeping across items; both instantiations
    (14-wide grid: two chunk-rows per tile; 28-wide: one), checked against the generic kernels."""
    from efficient_attention import _abi
    H, d, N = 3, 64, grid * grid
    dev = _dev()
    g = torch.Generator().manual_seed(11)
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.1).half().to(dev)
    bias = (0.5 * torch.randn(H, 49, 49, generator=g)).to(dev)
    ada = _abi_ada(_rand_ada(d, g), dev, 0.5)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(grid, grid), window=7, ext=0, chunk=chunk, chunk_ext=0)
    with _cluster_mode(cluster):
        out, path = _abi.eva_forward(q, k, v, geom, ada, bias=bias, return_path=True)
    assert path == (3 if cluster else 1)
    kb, bt = _abi.eva_chunk_stats(q, k, v, geom, ada)
    ref = _abi.eva_window_attention(q, k, v, geom, k_bar=kb, beta=bt, bias=bias)
    per_item = ((out.float() - ref.float()).view(B, N, H, d).pow(2).sum((1, 3)).sqrt() /
                ref.float().view(B, N, H, d).pow(2).sum((1, 3)).sqrt())
    assert float(per_item.max()) < TOL_F16, float(per_item.max())


class _causal_one_pass:
    """Run a block with the opt-in one-pass mode of the causal tcgen05 kernel (chunk statistics inside the window kernel)."""

    def __init__(self, on):
        self.on = on

    def __enter__(self):
        import ctypes
        from efficient_attention import _abi
        self.lib = _abi.load()
        self.lib.eva_debug_set_causal_one_pass.restype = ctypes.c_int
        self.prev = self.lib.eva_debug_set_causal_one_pass(ctypes.c_int(1 if self.on else 0))

    def __exit__(self, *exc):
        import ctypes
        self.lib.eva_debug_set_causal_one_pass(ctypes.c_int(self.prev))


@pytest.mark.parametrize('one_pass', [False, True])
@pytest.mark.parametrize('dtype,tol', [(torch.float16, TOL_F16), (torch.bfloat16, TOL_BF16)])
@pytest.mark.parametrize('chunk,with_noise,with_bias', [(256, False, False), (64, True, False), (128, False, True), (256, True, True)])
def test_causal_tcgen05_window_kernel_vs_oracle(chunk, with_noise, with_bias, dtype, tol, one_pass):
    """Causal EVA core (window 256, no halo, head_dim 64, 16-bit I/O: the c5 geometry) through the tcgen05 window kernel
    and the CTA-per-chunk statistics kernel, against the oracle on identical inputs.  path == 2 proves the tcgen05
    kernel ran (a silent fall-back to the CUDA-core kernel would hide a regression)."""
    from efficient_attention import _abi
    B, H, d, N, w = 2, 2, 64, 1024, 256
    g = torch.Generator().manual_seed(chunk + int(with_noise))
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.1).to(dtype)
    ada = _rand_ada(d, g)
    C = N // chunk
    noise = torch.randn(B, H, C, d, generator=g) if with_noise else None
    bias = None
    if with_bias:          # Toeplitz like the T5 bucketed bias: bias[i][j] = f(i - j) (entries with j > i are masked anyway)
        dist = torch.randn(w, generator=g)
        ii = torch.arange(w)
        bias = dist[(ii[:, None] - ii[None, :]).clamp(min=0)].unsqueeze(0)
    q64, k64, v64 = (qkv[:, :, i].permute(0, 2, 1, 3).double() for i in range(3))
    want, kbar_w, beta_w = O.eva_core(q64, k64, v64, seq_shape=(N,), window=w, ext=0, chunk=chunk, chunk_ext=0,
                                      **{k_: v_.double() for k_, v_ in ada.items()}, mu_coeff=1.0,
                                      noise=noise.double() if with_noise else None, causal=True, halo_right=False,
                                      mask_queries=True, bias=bias.double() if with_bias else None, return_stats=True)
    want = want.permute(0, 2, 1, 3).reshape(B, N, H * d)
    dev = _dev()
    qd = qkv.to(dev)
    q, k, v = qd[:, :, 0], qd[:, :, 1], qd[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(N,), window=w, ext=0, chunk=chunk, chunk_ext=0, causal=True, halo_left_only=True,
                             mask_queries=True, bias_toeplitz=with_bias)
    ada_s = _abi_ada(ada, dev, 1.0)
    nz = noise.to(dev) if with_noise else None
    kbar, beta = _abi.eva_chunk_stats(q, k, v, geom, ada_s, noise=nz)          # CTA-per-chunk kernel (chunk >= 64)
    assert rel_l2(kbar.cpu(), kbar_w) < 2e-5 and rel_l2(beta.cpu(), beta_w) < 2e-5
    with _causal_one_pass(one_pass):
        out, path = _abi.eva_forward(q, k, v, geom, ada_s, noise=nz, bias=bias.to(dev) if with_bias else None, return_path=True)
    assert path == 2
    if with_bias:          # without the caller's Toeplitz guarantee a biased call must stay on the CUDA-core kernel
        geom0 = _abi.eva_geometry(q, seq_shape=(N,), window=w, ext=0, chunk=chunk, chunk_ext=0, causal=True,
                                  halo_left_only=True, mask_queries=True)
        out0, path0 = _abi.eva_forward(q, k, v, geom0, ada_s, noise=nz, bias=bias.to(dev), return_path=True)
        assert path0 == 0 and rel_l2(out0.cpu(), want) < tol
    assert not torch.isnan(out).any()
    err = rel_l2(out.cpu(), want)
    assert err < tol, (chunk, with_noise, dtype, err)


def test_causal_tcgen05_many_windows_per_cta_fp16():
    """More windows than SMs (every CTA loops over several windows, both stages of the ring are reused) against the
    CUDA-core kernels fed with the same statistics."""
    import os
    import subprocess
    import sys
    from efficient_attention import _abi
    B, H, d, N = 6, 8, 64, 2048           # 6 * 8 * 8 = 384 windows
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.1).half().to(dev)
    ada = _abi_ada(_rand_ada(d, g), dev, 1.0)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(N,), window=256, ext=0, chunk=256, chunk_ext=0, causal=True, halo_left_only=True,
                             mask_queries=True)
    out, path = _abi.eva_forward(q, k, v, geom, ada, return_path=True)
    assert path == 2
    # reference: fp32 copies of the same 16-bit values take the CUDA-core kernels (the tcgen05 path is 16-bit only)
    q32, k32, v32 = (t.float() for t in (q, k, v))
    geom32 = _abi.eva_geometry(q32, seq_shape=(N,), window=256, ext=0, chunk=256, chunk_ext=0, causal=True,
                               halo_left_only=True, mask_queries=True)
    ref, path32 = _abi.eva_forward(q32, k32, v32, geom32, ada, return_path=True)
    assert path32 == 0
    per_item = ((out.float() - ref).view(B, N // 256, 256, H, d).pow(2).sum((2, 4)).sqrt() /
                ref.view(B, N // 256, 256, H, d).pow(2).sum((2, 4)).sqrt())
    assert float(per_item.max()) < TOL_F16, float(per_item.max())


def test_causal_time_major_views_fp16():
    """Batch / token strides exchanged ([T, B, ...] activations, as the causal module's fused q/k/v projection produces them):
    the tcgen05 window kernel builds its tensor maps with the two dimensions swapped; results must equal the batch-major call
    bit for bit (same kernels, same values, only the addressing differs)."""
    from efficient_attention import _abi
    B, H, d, N = 3, 4, 64, 1024
    dev = _dev()
    g = torch.Generator().manual_seed(11)
    tm = (torch.randn(N, B, 3, H, d, generator=g) * 1.1).half().to(dev)          # time-major storage
    bm = tm.transpose(0, 1).contiguous()                                         # batch-major copy of the same values
    ada = _abi_ada(_rand_ada(d, g), dev, 1.0)
    outs = []
    for src in (tm.transpose(0, 1), bm):
        q, k, v = src[:, :, 0], src[:, :, 1], src[:, :, 2]
        geom = _abi.eva_geometry(q, seq_shape=(N,), window=256, ext=0, chunk=128, chunk_ext=0, causal=True, halo_left_only=True,
                                 mask_queries=True)
        out, path = _abi.eva_forward(q, k, v, geom, ada, return_path=True)
        assert path == 2
        outs.append(out)
    assert torch.equal(outs[0], outs[1])


def test_causal_module_fused_projection_matches_separate_fp16():
    """CausalEVAttention under no_grad projects q, k, v with one GEMM on the [T, B, C] input (no transposed copies); with
    autograd enabled it takes the three separate projections.  Same module, same input: same output up to GEMM rounding."""
    import argparse
    import warnings
    import efficient_attention as ea
    dev = _dev()
    ns = argparse.Namespace(adaptive_proj='qk', num_chunks=None, chunk_size=128, causal=True, use_t5_rpe=True, window_size=256,
                            overlap_window=False)
    torch.manual_seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = ea.CausalEVAttention(embed_dim=256, num_heads=4, dropout=0.0, self_attention=True, attn_args=ns).to(dev).half().eval()
    x = torch.randn(512, 3, 256, device=dev, dtype=torch.float16)
    with torch.no_grad():
        y_fused = m(x, x, x, need_weights=False)[0]
        from efficient_attention import _abi
        assert 'qkv_fused' in _abi._MEMO[m]
        import copy
        copy.deepcopy(m)                       # the cache (ctypes structures) is not part of the module
    y_sep = m(x, x, x, need_weights=False)[0].detach()
    err = float((y_fused.float() - y_sep.float()).norm() / y_sep.float().norm())
    assert err < 1e-3, err
    # the cache follows in-place parameter updates
    with torch.no_grad():
        m.q_proj.weight.mul_(0.5)
        y2 = m(x, x, x, need_weights=False)[0]
    y2_sep = m(x, x, x, need_weights=False)[0].detach()
    assert float((y2.float() - y2_sep.float()).norm() / y2_sep.float().norm()) < 1e-3
    assert float((y2.float() - y_fused.float()).norm()) > 0


@pytest.mark.parametrize('dtype,tol', [(torch.float16, 2e-3), (torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize('proposal,with_noise', [('pool-mixed', False), ('pool', False), ('pool-mixed', True)])
def test_c4_lara_tcgen05_core_16bit(proposal, with_noise, dtype, tol):
    """BASELINE config c4 through the tcgen05 LARA core (mis-opt, one sample per landmark, 16-bit I/O) against the oracle
    evaluated in float64 on the same 16-bit weights and inputs (module level: the qkv / proj GEMMs round too, hence the
    looser tolerance).  The library's launch counter proves the tcgen05 core ran, not the CUDA-core kernels."""
    import ctypes
    import efficient_attention as ea
    from efficient_attention import _abi
    torch.manual_seed(14)
    m = ea.AttentionFactory.build_attention('lara', dict(dim=384, num_heads=6, num_landmarks=49,
                                                         proposal_gen=proposal, mis_type='mis-opt')).eval()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 2:
                p.normal_(0, 1.0 / math.sqrt(p.shape[1]))
    m = m.to(dtype)
    x = torch.randn(6, 14, 14, 384).to(dtype)
    cfg = dict(num_heads=6, num_landmarks=49, proposal_gen=proposal, mis_type='mis-opt', alpha_coeff=1.0)
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    noise = torch.randn(6, 6, 49, 64) if with_noise else None      # training-mode draw of lara.py:197-198 (one sample per landmark)
    want = O.lara_forward(sd, cfg, x.double(), noise=noise.double() if with_noise else None)
    lib = _abi.load()
    lib.eva_debug_lara_core_launches.restype = ctypes.c_int
    before = lib.eva_debug_lara_core_launches()
    with torch.no_grad():
        got = m.to(_dev())(x.to(_dev()), noise=noise.to(_dev()) if with_noise else None)
    assert lib.eva_debug_lara_core_launches() == before + 1
    assert not torch.isnan(got).any()
    err = rel_l2(got.cpu(), want)
    assert err < tol, (proposal, dtype, err)


def test_lara_tcgen05_core_many_items_fp16():
    """More (batch, head) items than SMs: every CTA loops over several items (both stages reused); checked against the
    CUDA-core kernels on fp32 copies of the same 16-bit values."""
    from efficient_attention import _abi
    dev = _dev()
    g = torch.Generator().manual_seed(21)
    B, H, d, N = 64, 6, 64, 196
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.1).half().to(dev)
    proj = _abi_ada(_rand_ada(d, g), dev, 1.0)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    kw = dict(seq_shape=(14, 14), landmarks=49, per_token_proj=False, mixed=1, mis_type='mis-opt', sample_mode=0,
              zero_padded=False, alpha_coeff=1.0, proj=proj)
    out = _abi.lara_forward(q, k, v, **kw)
    ref = _abi.lara_forward(q.float(), k.float(), v.float(), **kw)
    per_item = ((out.float() - ref).view(B, N, H, d).pow(2).sum((1, 3)).sqrt() / ref.view(B, N, H, d).pow(2).sum((1, 3)).sqrt())
    # landmarks, mixing and proposal statistics go through 16-bit MMA operands on this path: worst item of 384 within 1.5e-3
    assert float(per_item.max()) < 1.5e-3, float(per_item.max())
    assert float(per_item.mean()) < TOL_F16, float(per_item.mean())


def test_bench_line_has_the_contract_keys():
    """bench.py on a small batch: ONE JSON line carrying every key of the measurement contract (roofline, e2e with
    per-step copy bytes, launch count, clocks sampled during the timed region)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--steps', '3', '--warmup', '3', '--batch', '64', '--no-cpu', '--no-deit'],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'e2e', 'gpu_launches', 'clocks'):
        assert key in d, key
    assert d['steps'] == 3 and d['warmup'] == 3 and d['n_gpus'] == 1 and d['gpu_launches'] > 0
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    e = d['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] == 64 * 784 * 192 * 2 == e['d2h_bytes_per_step']
    assert e['value'] < d['value']                       # host copies are inside the e2e region
    assert d['config']['kernel_path'].startswith('fused tcgen05/TMA') and 'sm_mhz' in d['clocks']
    assert d['e2e']['copy_ceiling']['value'] > 0 and 0 < d['e2e']['frac_of_copy_ceiling'] < 1.2


# ---- round 2: parity gaps named by the round-1 review -------------------------------------------------
def _path_counts():
    import ctypes
    from efficient_attention import _abi
    lib = _abi.load()
    lib.eva_debug_path_count.restype = ctypes.c_int
    return [lib.eva_debug_path_count(ctypes.c_int(p)) for p in range(4)]


def _lara_core_launches():
    import ctypes
    from efficient_attention import _abi
    lib = _abi.load()
    lib.eva_debug_lara_core_launches.restype = ctypes.c_int
    return lib.eva_debug_lara_core_launches()


@pytest.mark.parametrize('mixed,with_noise', [(True, False), (False, False), (True, True)])
def test_lara_core_only_prequantised_fp16_vs_oracle(mixed, with_noise):
    """LARA CORE through the C ABI (no qkv / proj GEMMs): identical pre-quantised fp16 q/k/v on both sides, the oracle in float64,
    c4 geometry (14x14, 49 landmarks, mis-opt, one sample per landmark) -- the north_star bound of 1e-3 on the tcgen05 kernel
    itself (the module-level test above also carries the rounding of the two cuBLAS GEMMs)."""
    from efficient_attention import _abi
    B, H, d, gh = 8, 6, 64, 14
    N = gh * gh
    g = torch.Generator().manual_seed(31 + int(mixed) + 2 * int(with_noise))
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.1).half()
    p = _rand_ada(d, g)
    noise = torch.randn(B, H, 49, d, generator=g) if with_noise else None
    q64, k64, v64 = (qkv[:, :, i].permute(0, 2, 1, 3).double() for i in range(3))
    pd = {k_: v_.double() for k_, v_ in p.items()}
    q_bar, k_bar = O.lara_landmarks_2d(q64, k64, v64, gh, gh, 49, **pd, mixed=mixed, vmixed=False)
    want = O.lara_core(q64, k64, v64, q_bar, k_bar, mis_type='mis-opt', alpha_coeff=2.0,
                       noise=noise.double() if with_noise else None)
    want = want.permute(0, 2, 1, 3).reshape(B, N, H * d)
    dev = _dev()
    qd = qkv.to(dev)
    before = _lara_core_launches()
    out = _abi.lara_forward(qd[:, :, 0], qd[:, :, 1], qd[:, :, 2], seq_shape=(gh, gh), landmarks=49, per_token_proj=False,
                            mixed=int(mixed), mis_type='mis-opt', sample_mode=0, zero_padded=False, alpha_coeff=2.0,
                            proj=_abi_ada(p, dev, 1.0), noise=noise.to(dev) if with_noise else None)
    assert _lara_core_launches() == before + 1            # the tcgen05 core ran, not the CUDA-core kernels
    err = rel_l2(out.cpu(), want)
    per_item = ((out.cpu().double() - want).view(B, N, H, d).pow(2).sum((1, 3)).sqrt() / want.view(B, N, H, d).pow(2).sum((1, 3)).sqrt())
    assert err < TOL_F16, (mixed, with_noise, err)
    assert float(per_item.max()) < 1.5 * TOL_F16, float(per_item.max())


FAST_GOLDENS = {   # fixture -> the tcgen05 kernel family the 16-bit module must take (head_dim 64 geometries; `eva_2d_train` has
    # head_dim 32 and stays on the generic kernels, so the training-mode draws have fixtures of their own here)
    'eva_c1': 'eva', 'eva_c3_geom': 'eva', 'eva_c1_train': 'eva', 'eva_c3_train': 'eva',
    'lara_c4_geom': 'lara', 'lara_c4_train': 'lara', 'causal_c5_fast': 'causal',
}


@pytest.mark.parametrize('name', sorted(FAST_GOLDENS))
def test_fast_path_goldens_fp16(name):
    """The reference's own outputs (golden fixtures, float64 reference run) against the 16-bit tcgen05 kernels: the module is
    cast to fp16 and must take the fused path.  Two comparisons: against the oracle on the SAME fp16-rounded weights and input
    (kernel + GEMM rounding only; <= 2e-3 at module level) and against the fixture's reference output itself (adds the rounding
    of the weights and the input to 11 bits; <= 4e-3)."""
    cfg, sd, a = load_golden(name, dtype=torch.float32)
    m = build_module(cfg)
    m.load_state_dict(sd)
    m = m.half().to(_dev())
    m.train(a['noise'] is not None)
    sd16 = {k_: (v_.half().double() if v_.is_floating_point() else v_) for k_, v_ in sd.items()}
    x16 = a['x'].half()
    noise = a['noise'].double() if a['noise'] is not None else None
    if cfg['kind'] == 'eva':
        want16 = O.eva_forward(sd16, cfg, x16.double(), noise=noise)
    elif cfg['kind'] == 'causal_eva':
        want16 = O.causal_eva_forward(sd16, cfg, x16.double(), noise=noise)
    else:
        want16 = O.lara_forward(sd16, cfg, x16.double(), noise=noise)
    before, lara_before = _path_counts(), _lara_core_launches()
    y = run_module(m, cfg, a, _dev(), torch.float16)
    after, lara_after = _path_counts(), _lara_core_launches()
    if FAST_GOLDENS[name] == 'eva':
        assert after[0] == before[0] and (after[1] - before[1]) + (after[3] - before[3]) == 1, (before, after)
    elif FAST_GOLDENS[name] == 'causal':
        assert after[0] == before[0] and after[2] == before[2] + 1, (before, after)
    else:
        assert lara_after == lara_before + 1
    err16 = rel_l2(y.cpu(), want16)
    err_ref = rel_l2(y.cpu(), a['y'])
    assert err16 < 2e-3, (name, err16)
    assert err_ref < 4e-3, (name, err_ref)


@pytest.mark.parametrize('grid,chunk,B,cluster', [(14, 2, 400, False), (28, 4, 120, False), (28, 4, 200, True)])
def test_fused_kernel_many_items_vs_oracle_fp16(grid, chunk, B, cluster):
    """Same situation as test_fused_kernel_many_items_per_cta_fp16 (every CTA / cluster loops over several items: ring, barrier
    phase and buffer reuse across items) but judged by the float64 ORACLE, item by item, not by the repo's own generic kernels."""
    from efficient_attention import _abi
    H, d, N = 3, 64, grid * grid
    dev = _dev()
    g = torch.Generator().manual_seed(12)
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.1).half()
    bias = 0.5 * torch.randn(H, 49, 49, generator=g)
    ada = _rand_ada(d, g)
    qd = qkv.to(dev)
    q, k, v = qd[:, :, 0], qd[:, :, 1], qd[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(grid, grid), window=7, ext=0, chunk=chunk, chunk_ext=0)
    with _cluster_mode(cluster):
        out, path = _abi.eva_forward(q, k, v, geom, _abi_ada(ada, dev, 0.5), bias=bias.to(dev), return_path=True)
    assert path == (3 if cluster else 1)
    out = out.cpu().double().view(B, N, H, d)
    worst = 0.0
    for lo in range(0, B, 40):                                   # the oracle in batches of 40 images
        sl = slice(lo, min(B, lo + 40))
        q64, k64, v64 = (qkv[sl, :, i].permute(0, 2, 1, 3).double() for i in range(3))
        want = O.eva_core(q64, k64, v64, seq_shape=(grid, grid), window=7, ext=0, chunk=chunk, chunk_ext=0,
                          **{k_: v_.double() for k_, v_ in ada.items()}, mu_coeff=0.5, bias=bias.double())
        want = want.permute(0, 2, 1, 3)                          # [b, N, H, d]
        per_item = (out[sl] - want).pow(2).sum((1, 3)).sqrt() / want.pow(2).sum((1, 3)).sqrt()
        worst = max(worst, float(per_item.max()))
    assert worst < TOL_F16, worst


@pytest.mark.parametrize('one_pass', [False, True])
def test_c5_causal_core_full_shape_fp16_vs_oracle(one_pass):
    """BASELINE config c5 at FULL size through the tcgen05 causal kernels (two passes, and the opt-in one-pass mode): T=4096, h=8,
    d=64, window = chunk = 256, T5 bias, fp16, against the float64 oracle on identical pre-quantised q/k/v (path == 2)."""
    from efficient_attention import _abi
    B, H, d, N, w = 2, 8, 64, 4096, 256
    g = torch.Generator().manual_seed(41)
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.1).half()
    ada = _rand_ada(d, g)
    dist_tab = torch.randn(w, generator=g) * 0.5
    ii = torch.arange(w)
    bias = dist_tab[(ii[:, None] - ii[None, :]).clamp(min=0)].unsqueeze(0)
    q64, k64, v64 = (qkv[:, :, i].permute(0, 2, 1, 3).double() for i in range(3))
    want = O.eva_core(q64, k64, v64, seq_shape=(N,), window=w, ext=0, chunk=w, chunk_ext=0,
                      **{k_: v_.double() for k_, v_ in ada.items()}, mu_coeff=1.0, causal=True, halo_right=False,
                      mask_queries=True, bias=bias.double())
    want = want.permute(0, 2, 1, 3).reshape(B, N, H * d)
    dev = _dev()
    qd = qkv.to(dev)
    q, k, v = qd[:, :, 0], qd[:, :, 1], qd[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(N,), window=w, ext=0, chunk=w, chunk_ext=0, causal=True, halo_left_only=True,
                             mask_queries=True, bias_toeplitz=True)
    with _causal_one_pass(one_pass):
        out, path = _abi.eva_forward(q, k, v, geom, _abi_ada(ada, dev, 1.0), bias=bias.to(dev), return_path=True)
    assert path == 2
    err = rel_l2(out.cpu(), want)
    assert err < TOL_F16, err


def test_c5_causal_module_full_shape_fp16():
    """The c5 LAYER (CausalEVAttention, T=4096, B=2, C=512, h=8) in fp16 against the oracle on the same fp16-rounded weights;
    module level (four projections round too): <= 2e-3."""
    from argparse import Namespace
    import efficient_attention as ea
    torch.manual_seed(3)
    m = ea.CausalEVAttention(512, 8, self_attention=True, attn_args=Namespace(
        adaptive_proj='qk', num_chunks=None, chunk_size=256, causal=True, use_t5_rpe=True, window_size=256,
        overlap_window=False)).eval()
    with torch.no_grad():
        m.rel_pos_bias.relative_attention_bias.weight.normal_(0, 0.5)
    m = m.half()
    x = torch.randn(4096, 2, 512).half()
    cfg = dict(num_heads=8, window_size=256, overlap_window=False, chunk_size=256, num_chunks=None, causal=True,
               use_t5_rpe=True, adaptive_proj='qk')
    sd = {k_: v_.detach().double() for k_, v_ in m.state_dict().items()}
    want = O.causal_eva_forward(sd, cfg, x.double())
    before = _path_counts()
    with torch.no_grad():
        got = m.to(_dev())(x.to(_dev()), None, None)[0]
    after = _path_counts()
    assert after[2] == before[2] + 1 and after[0] == before[0]
    assert rel_l2(got.cpu(), want) < 2e-3


def test_causal_time_major_views_vs_oracle_fp16():
    """Time-major [T, B, ...] activations (tensor maps with batch / token exchanged) against the ORACLE, not only against the
    batch-major call of the same kernel."""
    from efficient_attention import _abi
    B, H, d, N = 3, 4, 64, 1024
    dev = _dev()
    g = torch.Generator().manual_seed(13)
    tm = (torch.randn(N, B, 3, H, d, generator=g) * 1.1).half()
    ada = _rand_ada(d, g)
    q64, k64, v64 = (tm[:, :, i].permute(1, 2, 0, 3).double() for i in range(3))          # [B, H, N, d]
    want = O.eva_core(q64, k64, v64, seq_shape=(N,), window=256, ext=0, chunk=128, chunk_ext=0,
                      **{k_: v_.double() for k_, v_ in ada.items()}, mu_coeff=1.0, causal=True, halo_right=False,
                      mask_queries=True)
    want = want.permute(0, 2, 1, 3).reshape(B, N, H * d)
    src = tm.to(dev).transpose(0, 1)
    q, k, v = src[:, :, 0], src[:, :, 1], src[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(N,), window=256, ext=0, chunk=128, chunk_ext=0, causal=True, halo_left_only=True,
                             mask_queries=True)
    out, path = _abi.eva_forward(q, k, v, geom, _abi_ada(ada, dev, 1.0), return_path=True)
    assert path == 2
    assert rel_l2(out.cpu(), want) < TOL_F16


def test_generic_kernels_accept_more_than_65535_batch_heads():
    """gridDim.y / .z stop at 65535; the generic window kernel folds (row block, window, batch * head) into gridDim.x
    (the reference has no such limit).  66 000 (batch, head) items of a tiny 1-D geometry, spot-checked against the oracle."""
    from efficient_attention import _abi
    B, H, d, N, w = 16500, 4, 16, 8, 4
    dev = _dev()
    g = torch.Generator().manual_seed(17)
    qkv = torch.randn(B, N, 3, H, d, generator=g)
    qd = qkv.to(dev)
    q, k, v = qd[:, :, 0], qd[:, :, 1], qd[:, :, 2]
    geom = _abi.eva_geometry(q, seq_shape=(N,), window=w, ext=0, chunk=0, chunk_ext=0)
    out = _abi.eva_window_attention(q, k, v, geom)
    sl = slice(B - 4, B)
    q64, k64, v64 = (qkv[sl, :, i].permute(0, 2, 1, 3).double() for i in range(3))
    want = O.local_core(q64, k64, v64, seq_shape=(N,), window=w, ext=0).permute(0, 2, 1, 3).reshape(4, N, H * d)
    assert rel_l2(out[sl].cpu(), want) < TOL_F32


def test_lara_core_many_items_vs_oracle_fp16():
    """More (batch, head) items than resident CTAs on the tcgen05 LARA core (stage / barrier reuse across items), judged item by
    item by the float64 oracle on the same pre-quantised fp16 q/k/v."""
    from efficient_attention import _abi
    dev = _dev()
    g = torch.Generator().manual_seed(22)
    B, H, d, gh = 64, 6, 64, 14
    N = gh * gh
    qkv = (torch.randn(B, N, 3, H, d, generator=g) * 1.1).half()
    p = _rand_ada(d, g)
    qd = qkv.to(dev)
    before = _lara_core_launches()
    out = _abi.lara_forward(qd[:, :, 0], qd[:, :, 1], qd[:, :, 2], seq_shape=(gh, gh), landmarks=49, per_token_proj=False,
                            mixed=1, mis_type='mis-opt', sample_mode=0, zero_padded=False, alpha_coeff=1.0,
                            proj=_abi_ada(p, dev, 1.0))
    assert _lara_core_launches() == before + 1
    out = out.cpu().double().view(B, N, H, d)
    pd = {k_: v_.double() for k_, v_ in p.items()}
    worst, mean = 0.0, 0.0
    for lo in range(0, B, 16):
        sl = slice(lo, lo + 16)
        q64, k64, v64 = (qkv[sl, :, i].permute(0, 2, 1, 3).double() for i in range(3))
        q_bar, k_bar = O.lara_landmarks_2d(q64, k64, v64, gh, gh, 49, **pd, mixed=True, vmixed=False)
        want = O.lara_core(q64, k64, v64, q_bar, k_bar, mis_type='mis-opt', alpha_coeff=1.0).permute(0, 2, 1, 3)
        per_item = (out[sl] - want).pow(2).sum((1, 3)).sqrt() / want.pow(2).sum((1, 3)).sqrt()
        worst = max(worst, float(per_item.max()))
        mean += float(per_item.sum()) / (B * H)
    assert mean < TOL_F16, mean
    assert worst < 1.5 * TOL_F16, worst          # worst of 384 items; the mean is the north_star figure


@pytest.mark.parametrize('seq_shape,window,ext,chunk,causal,with_mask,dtype', [
    ((14, 14), 7, 3, 2, False, False, torch.float16), ((28, 28), 7, 0, 4, False, False, torch.bfloat16),
    ((96,), 16, 8, 12, False, True, torch.float16), ((512,), 128, 128, 64, True, True, torch.float16),
    ((128,), 32, 0, 16, True, False, torch.bfloat16), ((14, 14), 7, 0, 0, False, False, torch.float16),
    ((197,), 197, 0, 0, False, True, torch.float16), ((784,), 784, 0, 0, False, False, torch.float16),
    ((640,), 320, 0, 64, True, False, torch.float16)])
def test_tcgen05_window_kernel_matches_the_cuda_core_kernel(seq_shape, window, ext, chunk, causal, with_mask, dtype):
    """eva_window_tc_sm100.cu (any geometry, head_dim 64, 16-bit) against window_attn_kernel of eva_generic.cu on identical
    inputs through `eva_window_attention`, and against the float64 oracle; the dispatch counter proves which kernel ran."""
    from efficient_attention import _abi
    dev = torch.device('cuda', 0)
    lib = _abi.load()
    B, H, d = 2, 3, 64
    N = math.prod(seq_shape)
    g = torch.Generator().manual_seed(N + window + ext + chunk)
    qkv = torch.randn(B, N, 3, H, d, generator=g).to(dtype)
    two_d = len(seq_shape) == 2
    L = window * window if two_d else window
    J = (window + 2 * ext) ** 2 if two_d else window + (ext if causal else 2 * ext)
    bias = 0.5 * torch.randn(H, L, J, generator=g) if L * J <= 1 << 16 else None
    mask = None
    if with_mask:
        mask = torch.zeros(B, N, dtype=torch.bool)
        mask[1, N - 9:] = True
    softmax_like = chunk == 0 and window == N
    geometry = dict(seq_shape=seq_shape, window=window, ext=ext, chunk=chunk, chunk_ext=0 if causal else ext, causal=causal,
                    halo_left_only=causal, mask_queries=causal, mask_is_neg_inf=softmax_like)
    qd = qkv.to(dev)
    q, k, v = qd[:, :, 0], qd[:, :, 1], qd[:, :, 2]
    geom = _abi.eva_geometry(q, **geometry)
    stats = {}
    ada = None
    if chunk:
        ada = _rand_ada(d, g)
        noise = torch.randn(B, H, _abi.num_chunks(geom), d, generator=g)
        kb, bt = _abi.eva_chunk_stats(q, k, v, geom, _abi_ada(ada, dev, 1.0 if causal else 0.5), pad_mask=None if mask is None else mask.to(dev),
                                      noise=noise.to(dev))
        stats = dict(k_bar=kb, beta=bt)

    def run(mode):
        lib.eva_debug_set_window_tc(mode)
        try:
            before = lib.eva_debug_window_tc_count()
            out = _abi.eva_window_attention(q, k, v, geom, pad_mask=None if mask is None else mask.to(dev),
                                            bias=None if bias is None else bias.to(dev), **stats)
            torch.cuda.synchronize()
            assert lib.eva_debug_window_tc_count() - before == mode
            return out.float().cpu()
        finally:
            lib.eva_debug_set_window_tc(-1)
    got, want = run(1), run(0)
    live = torch.ones(B, N, dtype=torch.bool) if (mask is None or not causal) else ~mask      # padded queries of the causal layer are don't-care rows
    assert torch.isfinite(got[live]).all()
    tol = 2e-3 if dtype == torch.float16 else 1.2e-2
    assert rel_l2(got[live], want[live]) < tol, rel_l2(got[live], want[live])


@pytest.mark.parametrize('seq_shape,chunk,chunk_ext,causal,with_mask,dtype', [
    ((14, 14), 2, 3, False, False, torch.float16), ((28, 28), 4, 3, False, False, torch.bfloat16), ((28, 28), 4, 0, False, False, torch.float16),
    ((96,), 12, 8, False, True, torch.float16), ((1024,), 16, 32, False, True, torch.float16), ((128,), 16, 0, True, True, torch.bfloat16),
    ((90,), 45, 40, False, True, torch.float16)])
def test_fast_chunk_statistics_kernel_matches_the_generic_one(seq_shape, chunk, chunk_ext, causal, with_mask, dtype):
    """chunk_stats_fast_kernel (head_dim 64, 16-bit I/O, <= 128 slots per chunk) against chunk_stats_kernel: the same 16-bit values
    handed over as float32 take the generic kernel; both compute in float32."""
    from efficient_attention import _abi
    dev = torch.device('cuda', 0)
    B, H, d = 3, 2, 64
    N = math.prod(seq_shape)
    g = torch.Generator().manual_seed(N + chunk + chunk_ext)
    qkv = torch.randn(B, N, 3, H, d, generator=g).to(dtype).to(dev)
    mask = None
    if with_mask:
        mask = torch.zeros(B, N, dtype=torch.bool, device=dev)
        mask[1, N - 9:] = True
        mask[2, :5] = True
    ada = _abi_ada(_rand_ada(d, g), dev, 1.0 if causal else 0.5)
    window = chunk * 2 if N % (chunk * 2) == 0 else chunk
    geometry = dict(seq_shape=seq_shape, window=window if len(seq_shape) == 1 else 7, ext=0, chunk=chunk, chunk_ext=chunk_ext, causal=causal,
                    halo_left_only=causal)
    noise = None

    def stats(x):
        q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
        geom = _abi.eva_geometry(q, **geometry)
        nz = torch.randn(B, H, _abi.num_chunks(geom), d, generator=torch.Generator().manual_seed(1)).to(dev)
        return _abi.eva_chunk_stats(q, k, v, geom, ada, pad_mask=mask, noise=nz)
    (kb, bt), (kb32, bt32) = stats(qkv), stats(qkv.float())
    assert rel_l2(kb.cpu(), kb32.cpu()) < 1e-5 and rel_l2(bt.cpu(), bt32.cpu()) < 1e-5, (rel_l2(kb.cpu(), kb32.cpu()), rel_l2(bt.cpu(), bt32.cpu()))
