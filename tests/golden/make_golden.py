"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Run only in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference package imports ``timm.models.layers.trunc_normal_``; timm is not installed and there
is no network, so a 3-line shim is written to a temp dir and put on sys.path.
Each fixture holds: the config, the module state_dict (float32), the input (float32), optional
padding mask / noise, and the reference output computed with the module in float64 on those values.
The reference ships no golden vectors of its own for this path, so these ARE the pins.
"""
import argparse
import json
import math
import os
import sys
import tempfile
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/efficient-attention'


def _import_reference():
    shim = tempfile.mkdtemp(prefix='timm_shim_')
    os.makedirs(os.path.join(shim, 'timm', 'models'))
    open(os.path.join(shim, 'timm', '__init__.py'), 'w').close()
    open(os.path.join(shim, 'timm', 'models', '__init__.py'), 'w').close()
    with open(os.path.join(shim, 'timm', 'models', 'layers.py'), 'w') as f:
        f.write('from torch.nn.init import trunc_normal_\n')
    sys.path.insert(0, shim)
    sys.path.insert(0, REF)
    import efficient_attention as ref  # noqa
    assert ref.__file__.startswith(REF), ref.__file__
    return ref


def _lively_init(module, seed):
    """Replace the reference's near-zero init (std .02 => uniform softmax everywhere) by weights that
    give logits of order 1, so that masking / bias / softmax mistakes are visible in the output."""
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        with torch.no_grad():
            if name.endswith('relative_attention_bias.weight') or 'bias_table' in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.7)
            elif p.dim() == 3:                       # learnable random-feature projections keep their orthogonal init
                continue
            elif p.dim() == 2:
                p.copy_(torch.randn(p.shape, generator=g) * (1.3 / math.sqrt(p.shape[1])))
            elif name.endswith('.weight'):           # LayerNorm gain
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            else:                                    # any bias
                p.copy_(0.2 * torch.randn(p.shape, generator=g))


def _save(name, cfg, module, arrays):
    sd = {k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}
    out = {f'sd::{k}': v for k, v in sd.items()}
    out.update({k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in arrays.items() if v is not None})
    out['cfg'] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name:28s} {os.path.getsize(path) / 1024:8.1f} KiB  y {tuple(arrays["y"].shape)}')


def _run(module, fn, train_seed=None):
    """Run in float64; in training mode seed the global RNG so the noise can be re-derived."""
    module = module.double()
    if train_seed is None:
        module.eval()
    else:
        module.train()
        torch.manual_seed(train_seed)
    with torch.no_grad():
        return fn(module)


def _noise(seed, shape):
    torch.manual_seed(seed)
    return torch.randn(shape, dtype=torch.float64)


def gen_eva(ref, name, *, B, shape, dim, heads, window, landmarks, attn_2d, overlap=False, use_rpe=False,
            use_t5=False, adaptive='default', mask_tail=None, train_seed=None, seed=0):
    cfg = dict(kind='eva', dim=dim, num_heads=heads, window_size=window, num_landmarks=landmarks, attn_2d=attn_2d,
               overlap_window=overlap, use_rpe=use_rpe, use_t5_rpe=use_t5, adaptive_proj=adaptive, qkv_bias=True)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = ref.AttentionFactory.build_attention('eva', dict(
            dim=dim, num_heads=heads, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False, use_rpe=use_rpe,
            window_size=window, attn_2d=attn_2d, overlap_window=overlap, adaptive_proj=adaptive,
            num_landmarks=landmarks, use_t5_rpe=use_t5))
    _lively_init(m, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn((B,) + tuple(shape) + (dim,), generator=g)
    mask = None
    if mask_tail is not None:
        n = int(np.prod(shape))
        mask = torch.zeros(B, n, dtype=torch.bool)
        for b, t in enumerate(mask_tail):
            if t:
                mask[b, n - t:] = True
    y = _run(m, lambda mod: mod(x.double(), mask), train_seed)
    noise = None
    if train_seed is not None:
        n_pad = int(np.prod(shape)) if attn_2d else int(math.ceil(np.prod(shape) / window) * window)
        chunk = int(math.sqrt(n_pad // landmarks)) if attn_2d else n_pad // landmarks
        ext = max(1, window // 2) if overlap else 0
        n_chunks = (shape[0] // chunk) * (shape[1] // chunk) if attn_2d else n_pad // chunk
        noise = _noise(train_seed, (B, heads, n_chunks, dim // heads))
    _save(name, cfg, m.float(), dict(x=x, mask=mask, noise=noise, y=y))


def gen_local(ref, name, *, kind, B, shape, dim, heads, window=4, attn_2d=False, overlap=False, use_rpe=False,
              mask_tail=None, seed=0):
    cfg = dict(kind=kind, dim=dim, num_heads=heads, window_size=window, attn_2d=attn_2d, overlap_window=overlap,
               use_rpe=use_rpe, qkv_bias=True)
    args = dict(dim=dim, num_heads=heads, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False)
    if kind == 'local':
        args.update(use_rpe=use_rpe, window_size=window, attn_2d=attn_2d, overlap_window=overlap)
    m = ref.AttentionFactory.build_attention(kind, args)
    _lively_init(m, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn((B,) + tuple(shape) + (dim,), generator=g)
    mask = None
    if mask_tail is not None:
        n = int(np.prod(shape))
        mask = torch.zeros(B, n, dtype=torch.bool)
        for b, t in enumerate(mask_tail):
            if t:
                mask[b, n - t:] = True
    y = _run(m, lambda mod: mod(x.double(), mask))
    _save(name, cfg, m.float(), dict(x=x, mask=mask, y=y))


def gen_lara(ref, name, *, B, shape, dim, heads, landmarks, proposal_gen, mis_type='mis-opt', alpha=1.0,
             antithetic=False, multisample=False, mask_tail=None, train_seed=None, seed=0, pool='light'):
    cfg = dict(kind='lara', dim=dim, num_heads=heads, num_landmarks=landmarks, proposal_gen=proposal_gen,
               mis_type=mis_type, alpha_coeff=alpha, use_antithetics=antithetic, use_multisample=multisample,
               pool_module_type=pool, qkv_bias=True)
    m = ref.AttentionFactory.build_attention('lara', dict(
        dim=dim, num_heads=heads, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False, num_landmarks=landmarks,
        kernel_size=None, proposal_gen=proposal_gen, use_antithetics=antithetic, use_multisample=multisample,
        pool_module_type=pool, mis_type=mis_type, alpha_coeff=alpha))
    _lively_init(m, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn((B,) + tuple(shape) + (dim,), generator=g)
    mask = None
    if mask_tail is not None:
        n = int(np.prod(shape))
        mask = torch.zeros(B, n, dtype=torch.bool)
        for b, t in enumerate(mask_tail):
            if t:
                mask[b, n - t:] = True
    y = _run(m, lambda mod: mod(x.double(), mask), train_seed)
    noise = None
    if train_seed is not None:
        n = int(np.prod(shape))
        s = min(landmarks, n) if len(shape) == 1 else int(math.sqrt(landmarks)) ** 2
        noise = _noise(train_seed, (B, heads, 2 * s if multisample else s, dim // heads))
    _save(name, cfg, m.float(), dict(x=x, mask=mask, noise=noise, y=y))


def gen_causal(ref, name, *, T, B, dim, heads, window, chunk_size=None, num_chunks=None, causal=True, use_t5=True,
               overlap=False, adaptive='qk', mask_tail=None, train_seed=None, seed=0):
    cfg = dict(kind='causal_eva', embed_dim=dim, num_heads=heads, window_size=window, chunk_size=chunk_size,
               num_chunks=num_chunks, causal=causal, use_t5_rpe=use_t5, overlap_window=overlap, adaptive_proj=adaptive)
    m = ref.CausalEVAttention(embed_dim=dim, num_heads=heads, self_attention=True, attn_args=Namespace(
        adaptive_proj=adaptive, num_chunks=num_chunks, chunk_size=chunk_size, causal=causal, use_t5_rpe=use_t5,
        window_size=window, overlap_window=overlap))
    _lively_init(m, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(T, B, dim, generator=g)
    mask = None
    if mask_tail is not None:
        mask = torch.zeros(B, T, dtype=torch.bool)
        for b, t in enumerate(mask_tail):
            if t:
                mask[b, T - t:] = True
    y = _run(m, lambda mod: mod(x.double(), x.double(), x.double(), key_padding_mask=mask)[0], train_seed)
    noise = None
    if train_seed is not None:
        n_pad = int(math.ceil(T / window) * window)
        cs = chunk_size if chunk_size is not None else n_pad // num_chunks
        noise = _noise(train_seed, (B, heads, n_pad // cs, dim // heads))
    _save(name, cfg, m.float(), dict(x=x, mask=mask, noise=noise, y=y))


def _tail_mask(B, n, mask_tail):
    if mask_tail is None:
        return None
    mask = torch.zeros(B, n, dtype=torch.bool)
    for b, t in enumerate(mask_tail):
        if t:
            mask[b, n - t:] = True
    return mask


def gen_performer(ref, name, *, B, shape, dim, heads, method='favorp', approx=64, cos=False, scheme='default', mask_tail=None,
                  train_seed=None, seed=0):
    """kernelized_attention.py.  Training mode with sample_scheme 'default' draws a fresh Gaussian projection with torch.randn as
    the first use of the global RNG in forward: it is re-derived from the seed and stored as `proj`."""
    cfg = dict(kind='performer', dim=dim, num_heads=heads, proj_method=method, approx_attn_dim=approx, cos_weighting=cos,
               sample_scheme=scheme, qkv_bias=True)
    m = ref.AttentionFactory.build_attention('performer', dict(
        dim=dim, num_heads=heads, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False, approx_attn_dim=approx,
        proj_method=method, cos_weighting=cos, sample_scheme=scheme))
    _lively_init(m, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn((B,) + tuple(shape) + (dim,), generator=g)
    mask = _tail_mask(B, int(np.prod(shape)), mask_tail)
    y = _run(m, lambda mod: mod(x.double(), mask), train_seed)
    proj = None
    if train_seed is not None and scheme == 'default' and method in ('favorp', 'relu', 'fourier'):
        torch.manual_seed(train_seed)
        proj = torch.randn(heads, approx, dim // heads, dtype=torch.float64)
    _save(name, cfg, m.float(), dict(x=x, mask=mask, proj=proj, y=y))


def gen_ra(ref, name, *, B, shape, dim, heads, num_samples, train_seed=None, eval_seed=1234, seed=0):
    """randomized_attention.py.  The key index drawn with torch.multinomial (and the Gaussian noise of training mode, drawn after
    it) are re-derived from the seed with the module's own q, k and stored as `k_ind` / `noise`."""
    cfg = dict(kind='ra', dim=dim, num_heads=heads, num_samples=num_samples, qkv_bias=True)
    m = ref.AttentionFactory.build_attention('ra', dict(
        dim=dim, num_heads=heads, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False, num_samples=num_samples))
    _lively_init(m, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn((B,) + tuple(shape) + (dim,), generator=g)
    m = m.double()
    m.train(train_seed is not None)
    rng_seed = train_seed if train_seed is not None else eval_seed
    with torch.no_grad():
        torch.manual_seed(rng_seed)
        y = m(x.double(), None)
        q, k, v = m.proj_and_split_heads(x.double())
        b, h, n, d = q.shape
        torch.manual_seed(rng_seed)
        k_ind = None
        mu_like = q + k                      # same memory layout as the reference's mu (randn_like fills in memory order)
        if num_samples not in (0, -1):
            pi = torch.softmax(torch.einsum('...nd,...md->...nm', m.scale * q, k), dim=-1)
            k_ind = torch.multinomial(pi.reshape(b * h * n, n), 1, replacement=True).reshape(b, h, n)
            mu_like = q + torch.gather(k, 2, k_ind.unsqueeze(-1).expand(-1, -1, -1, d))
        noise = torch.randn_like(mu_like).contiguous() if train_seed is not None else None
    _save(name, cfg, m.float(), dict(x=x, mask=None, k_ind=k_ind, noise=noise, y=y))


def gen_scatterbrain(ref, name, *, B, shape, dim, heads, window, attn_2d, overlap=False, use_rpe=False, approx=64,
                     mask_tail=None, train_seed=None, seed=0):
    cfg = dict(kind='scatterbrain', dim=dim, num_heads=heads, window_size=window, attn_2d=attn_2d, overlap_window=overlap,
               use_rpe=use_rpe, approx_attn_dim=approx, proj_method='favorp', qkv_bias=True)
    m = ref.AttentionFactory.build_attention('scatterbrain', dict(
        dim=dim, num_heads=heads, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False, use_rpe=use_rpe, window_size=window,
        attn_2d=attn_2d, overlap_window=overlap, approx_attn_dim=approx))
    _lively_init(m, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn((B,) + tuple(shape) + (dim,), generator=g)
    mask = _tail_mask(B, int(np.prod(shape)), mask_tail)
    y = _run(m, lambda mod: mod(x.double(), mask), train_seed)
    proj = None
    if train_seed is not None:
        torch.manual_seed(train_seed)
        proj = torch.randn(heads, approx, dim // heads, dtype=torch.float64)
    _save(name, cfg, m.float(), dict(x=x, mask=mask, proj=proj, y=y))


def main_rfa(ref):
    # --- Performer (kernelized_attention.py) ---
    gen_performer(ref, 'perf_favorp_2d', B=2, shape=(14, 14), dim=128, heads=2, seed=81)
    gen_performer(ref, 'perf_favorp_mask', B=3, shape=(45,), dim=64, heads=2, mask_tail=[0, 7, 20], seed=83)
    gen_performer(ref, 'perf_favorp_train', B=2, shape=(40,), dim=128, heads=2, train_seed=85, seed=84)
    gen_performer(ref, 'perf_favorp_cos', B=2, shape=(33,), dim=64, heads=2, cos=True, approx=32, seed=86)
    gen_performer(ref, 'perf_relu_fixed', B=2, shape=(8, 8), dim=64, heads=2, method='relu', scheme='fixed', seed=87)
    gen_performer(ref, 'perf_fourier_learn', B=2, shape=(31,), dim=64, heads=4, method='fourier', scheme='learnable', approx=24, mask_tail=[0, 4], seed=89)
    gen_performer(ref, 'perf_dpfp', B=2, shape=(29,), dim=64, heads=2, method='dpfp', approx=128, seed=91)
    gen_performer(ref, 'perf_relu_only_cos', B=1, shape=(50,), dim=64, heads=2, method='relu-only', cos=True, seed=93)
    gen_performer(ref, 'perf_sigmoid_only', B=2, shape=(6, 6), dim=128, heads=2, method='sigmoid-only', mask_tail=[0, 5], seed=95)
    gen_performer(ref, 'perf_mlp_fourier', B=2, shape=(27,), dim=64, heads=2, method='mlp-fourier', approx=64, seed=97)
    # --- randomized attention (randomized_attention.py) ---
    gen_ra(ref, 'ra_mean', B=2, shape=(37,), dim=64, heads=2, num_samples=0, seed=101)
    gen_ra(ref, 'ra_expect_2d', B=1, shape=(8, 8), dim=128, heads=2, num_samples=-1, seed=103)
    gen_ra(ref, 'ra_sample', B=2, shape=(41,), dim=64, heads=2, num_samples=1, seed=105)
    gen_ra(ref, 'ra_sample_train', B=2, shape=(14, 14), dim=128, heads=2, num_samples=1, train_seed=109, seed=107)
    # --- ScatterBrain (scatterbrain_attention.py) ---
    gen_scatterbrain(ref, 'sb_2d_rpe', B=2, shape=(14, 14), dim=128, heads=2, window=7, attn_2d=True, use_rpe=True, seed=111)
    # overlap_window=True is not pinned: the reference returns NaN there (the zero-padded halo slots carry log-feature 0, which
    # makes the "local" log-sum-exp exceed the global one; log_add_exp(.., mask=(1, -1)) then takes the log of a negative number)
    gen_scatterbrain(ref, 'sb_2d_small', B=1, shape=(8, 8), dim=64, heads=2, window=4, attn_2d=True, use_rpe=True, approx=32, seed=113)
    gen_scatterbrain(ref, 'sb_1d_pad_mask', B=3, shape=(45,), dim=64, heads=2, window=8, attn_2d=False, use_rpe=True, mask_tail=[0, 6, 19], seed=115)
    gen_scatterbrain(ref, 'sb_1d_norpe', B=2, shape=(48,), dim=64, heads=4, window=16, attn_2d=False, mask_tail=[3, 0], seed=117)
    gen_scatterbrain(ref, 'sb_2d_train', B=2, shape=(14, 14), dim=128, heads=2, window=7, attn_2d=True, use_rpe=True, train_seed=121, seed=119)


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument('--only-rfa', action='store_true', help='only the performer / ra / scatterbrain fixtures')
    args = ap.parse_args()
    ref = _import_reference()
    torch.set_num_threads(8)
    main_rfa(ref)
    if args.only_rfa:
        return
    # --- EVA (eva.py) ---
    # BASELINE config c1, exact shape
    gen_eva(ref, 'eva_c1', B=2, shape=(14, 14), dim=192, heads=3, window=7, landmarks=49, attn_2d=True, use_rpe=True)
    # c3 geometry (28x28, chunk 4x4) at reduced width
    gen_eva(ref, 'eva_c3_geom', B=1, shape=(28, 28), dim=128, heads=2, window=7, landmarks=49, attn_2d=True, use_rpe=True, seed=3)
    gen_eva(ref, 'eva_2d_overlap', B=2, shape=(14, 14), dim=64, heads=2, window=7, landmarks=49, attn_2d=True, overlap=True, use_rpe=True, seed=5)
    gen_eva(ref, 'eva_2d_train', B=2, shape=(14, 14), dim=64, heads=2, window=7, landmarks=49, attn_2d=True, use_rpe=True, train_seed=11, seed=7)
    # training-mode draws on the geometries the tcgen05 kernels take (head_dim 64): c3 and c1 grids
    gen_eva(ref, 'eva_c3_train', B=1, shape=(28, 28), dim=128, heads=2, window=7, landmarks=49, attn_2d=True, use_rpe=True, train_seed=63, seed=61)
    gen_eva(ref, 'eva_c1_train', B=2, shape=(14, 14), dim=128, heads=2, window=7, landmarks=49, attn_2d=True, use_rpe=True, train_seed=67, seed=65)
    gen_eva(ref, 'eva_2d_noln', B=1, shape=(8, 8), dim=64, heads=2, window=4, landmarks=16, attn_2d=True, adaptive='no-ln', seed=9)
    gen_eva(ref, 'eva_2d_none', B=1, shape=(8, 8), dim=64, heads=2, window=4, landmarks=16, attn_2d=True, adaptive='none', seed=13)
    gen_eva(ref, 'eva_2d_t5', B=1, shape=(14, 14), dim=64, heads=2, window=7, landmarks=49, attn_2d=True, use_t5=True, seed=14)
    gen_eva(ref, 'eva_1d_t5_mask', B=3, shape=(50,), dim=64, heads=2, window=8, landmarks=7, attn_2d=False, overlap=True, use_t5=True, mask_tail=[0, 5, 17], seed=15)
    gen_eva(ref, 'eva_1d_rpe', B=2, shape=(48,), dim=64, heads=4, window=8, landmarks=6, attn_2d=False, use_rpe=True, mask_tail=[3, 0], seed=17)
    gen_eva(ref, 'eva_1d_train', B=2, shape=(45,), dim=64, heads=2, window=6, landmarks=8, attn_2d=False, overlap=True, train_seed=21, seed=19)
    # --- softmax / local (abstract_attention.py, local_attention.py) ---
    gen_local(ref, 'softmax_mask', kind='softmax', B=2, shape=(37,), dim=64, heads=2, mask_tail=[0, 9], seed=23)
    gen_local(ref, 'local_2d_rpe', kind='local', B=2, shape=(8, 8), dim=64, heads=2, window=4, attn_2d=True, overlap=True, use_rpe=True, seed=25)
    gen_local(ref, 'local_1d_mask', kind='local', B=2, shape=(30,), dim=64, heads=2, window=8, overlap=True, use_rpe=True, mask_tail=[4, 0], seed=27)
    # --- LARA (lara.py) ---
    gen_lara(ref, 'lara_c4_geom', B=2, shape=(14, 14), dim=128, heads=2, landmarks=49, proposal_gen='pool-mixed', seed=29)
    gen_lara(ref, 'lara_c4_train', B=2, shape=(14, 14), dim=128, heads=2, landmarks=49, proposal_gen='pool-mixed', train_seed=71, seed=69)
    # pool_module_type == 'dense' (Linear / LayerNorm across heads, lara.py:36-39, 131-139; the PVT configuration of the README)
    gen_lara(ref, 'lara_2d_dense', B=2, shape=(14, 14), dim=128, heads=2, landmarks=49, proposal_gen='pool-mixed', pool='dense', alpha=2.0, seed=75)
    gen_lara(ref, 'lara_2d_dense_vmixed', B=1, shape=(8, 8), dim=64, heads=2, landmarks=16, proposal_gen='pool-vmixed', pool='dense', mis_type='mis-bh', seed=77)
    gen_lara(ref, 'lara_2d_pool_bh', B=1, shape=(10, 12), dim=64, heads=2, landmarks=16, proposal_gen='pool', mis_type='mis-bh', seed=31)
    gen_lara(ref, 'lara_2d_vmixed_biased', B=1, shape=(8, 8), dim=64, heads=2, landmarks=16, proposal_gen='pool-vmixed', mis_type='mis-biased', seed=33)
    gen_lara(ref, 'lara_2d_noparam', B=1, shape=(8, 8), dim=64, heads=2, landmarks=16, proposal_gen='no-param-pool', alpha=0.5, seed=35)
    gen_lara(ref, 'lara_2d_train_anti', B=1, shape=(14, 14), dim=64, heads=2, landmarks=49, proposal_gen='pool-mixed', antithetic=True, train_seed=41, seed=37)
    gen_lara(ref, 'lara_2d_train_multi', B=1, shape=(8, 8), dim=64, heads=2, landmarks=16, proposal_gen='pool', multisample=True, train_seed=43, seed=39)
    gen_lara(ref, 'lara_1d_uneven_mask', B=2, shape=(61,), dim=64, heads=2, landmarks=8, proposal_gen='adaptive-1d', mask_tail=[0, 6], seed=45)
    gen_lara(ref, 'lara_1d_even', B=2, shape=(64,), dim=64, heads=2, landmarks=8, proposal_gen='adaptive-1d', mis_type='mis-bh', train_seed=47, seed=49)
    # --- causal EVA (causal_eva.py) ---
    # the reference's own self-check configuration (causal_eva.py:916-950), shortened sequence
    gen_causal(ref, 'causal_selfcheck', T=128, B=2, dim=128, heads=8, window=64, chunk_size=16, seed=51)
    gen_causal(ref, 'causal_c5_geom', T=96, B=2, dim=128, heads=2, window=32, chunk_size=32, seed=53)
    # window = chunk = 256, head_dim 64: the geometry of the tcgen05 causal kernels (c5) at two windows
    gen_causal(ref, 'causal_c5_fast', T=512, B=1, dim=128, heads=2, window=256, chunk_size=256, seed=73)
    gen_causal(ref, 'causal_overlap_mask', T=75, B=3, dim=64, heads=2, window=16, chunk_size=8, overlap=True, mask_tail=[0, 7, 30], seed=55)
    gen_causal(ref, 'causal_numchunks_train', T=64, B=2, dim=64, heads=2, window=16, num_chunks=8, use_t5=False, adaptive='no-ln', train_seed=61, seed=57)
    gen_causal(ref, 'noncausal_flag', T=64, B=1, dim=64, heads=2, window=16, chunk_size=8, causal=False, overlap=True, seed=59)


if __name__ == '__main__':
    main()
