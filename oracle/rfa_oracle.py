"""CPU oracle for the random-feature modules: Performer ('performer'), randomized attention ('ra'), ScatterBrain ('scatterbrain').

*** TEST INFRASTRUCTURE ONLY. ***  Same rules as ``oracle/eva_oracle.py``: nothing in the product package imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may.

An independent restatement (plain torch CPU arithmetic, normally float64) of

    efficient-attention/efficient_attention/kernelized_attention.py:12-127,185-330   (feature maps, linear attention, KernelizedAttention)
    efficient-attention/efficient_attention/randomized_attention.py:24-55            (RandomizedAttention._apply_attention)
    efficient-attention/efficient_attention/scatterbrain_attention.py:10-164         (log-FAVOR+ features, ScatterBrain.forward)
    efficient-attention/efficient_attention/attn_utils.py:44-51                      (log_add_exp)

Windows are explicit gather tables (``eva_oracle.group_index_*``) instead of ``F.pad`` + ``as_strided``; the random draws of the
reference (projection matrix in training mode, the multinomial key sample of 'ra', its Gaussian noise) are explicit arguments.

Pinning: the reference ships no golden vectors for these modules; the oracle is pinned against outputs of the reference itself
(``tests/golden/make_golden.py`` -> ``tests/golden/perf_*.npz``, ``ra_*.npz``, ``sb_*.npz``; ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math

import torch

from .eva_oracle import (_pad_tokens, _split_qkv, group_index_1d, group_index_2d, linear, local_bias_from_state, take_groups,
                         take_mask_groups)


# --------------------------------------------------------------------------------------------
# feature maps (kernelized_attention.py:12-113, 155-178)
# --------------------------------------------------------------------------------------------
def feature_dim(method: str, approx_dim: int, head_dim: int) -> int:
    if method == 'fourier':
        return 2 * approx_dim
    if method == 'dpfp':
        return 2 * head_dim * ((approx_dim // head_dim) // 2)
    if method in ('relu-only', 'sigmoid-only'):
        return head_dim
    return approx_dim


def features(x, method: str, is_query: bool, proj=None, nu: int = 1, mlp_w=None, mlp_b=None):
    """x [B, h, N, d] -> phi(x) [B, h, N, M]; proj [h, m, d]."""
    d = x.shape[-1]
    dn = d ** -0.25
    if method in ('favorp', 'relu', 'fourier'):
        m = proj.shape[1]
        dd = torch.einsum('bhnd,hjd->bhnj', dn * x, proj)
        half_sq = 0.5 * dn * dn * (x * x).sum(-1, keepdim=True)
        if method == 'favorp':                                   # :21-55
            stab = (dd.amax(-1, keepdim=True) if is_query else dd.amax((-1, -2), keepdim=True)).detach()   # detached in the reference too
            return m ** -0.5 * torch.exp(dd - half_sq - stab) + 1e-4
        if method == 'relu':                                     # :92-113 (generalized_projection, eps = 1e-3)
            return torch.relu(m ** -0.5 * dd) + 1e-3
        hq = torch.exp(half_sq - half_sq.amax(-2, keepdim=True).detach())  # :57-87: the max runs over the TOKENS
        return hq * (m ** -0.5) * torch.cat([torch.sin(dd), torch.cos(dd)], -1)
    if method == 'dpfp':                                         # :12-19
        x2 = torch.cat([torch.relu(x), torch.relu(-x)], -1)
        rolled = torch.cat([x2.roll(shifts=j, dims=-1) for j in range(1, nu + 1)], -1)
        return torch.cat([x2] * nu, -1) * rolled
    if method == 'relu-only':                                    # :89-90 (eps = 0.1)
        return torch.relu(x) + 0.1
    if method == 'sigmoid-only':
        return torch.sigmoid(x) + 0.1
    if method == 'mlp-fourier':                                  # :155-178
        px = torch.einsum('bhnd,hjd->bhnj', x, proj)
        ff = torch.cat([px.cos(), px.sin()], -1) * d ** -0.5
        return torch.relu(linear(ff, mlp_w, mlp_b))
    raise KeyError(method)


def linear_attention(qf, kf, v, cos_weighting: bool = False):
    """kernelized_attention.py:115-153: phi(q) (phi(k)^T v) / clamp(phi(q) . sum phi(k), 1e-2); cosFormer re-weighting = the same
    with the features doubled to [phi cos(pi n / 2N) ; phi sin(pi n / 2N)]."""
    if cos_weighting:
        n = v.shape[-2]
        ang = (math.pi / 2) * torch.arange(n, dtype=v.dtype) / n
        c, s = torch.cos(ang).view(1, 1, n, 1), torch.sin(ang).view(1, 1, n, 1)
        qf, kf = torch.cat([qf * c, qf * s], -1), torch.cat([kf * c, kf * s], -1)
    kv = kf.transpose(-1, -2) @ v
    den = (qf * kf.sum(-2, keepdim=True)).sum(-1, keepdim=True)
    return (qf @ kv) / den.clamp(min=1e-2)


def performer_core(q, k, v, *, method, proj=None, nu=1, mlp_w=None, mlp_b=None, cos_weighting=False, pad_mask=None,
                   f32_linear=False):
    """KernelizedAttention._apply_attention (:301-320): padded keys' features are zeroed AFTER the (global) stabilisers.
    f32_linear: the reference casts the features and v to float32 for the linear attention whatever the module's dtype (:319),
    so its "float64" outputs carry float32 rounding from that step; True reproduces the cast."""
    qf = features(q, method, True, proj, nu, mlp_w, mlp_b)
    kf = features(k, method, False, proj, nu, mlp_w, mlp_b)
    if pad_mask is not None:
        kf = kf.masked_fill(pad_mask.bool().unsqueeze(1).unsqueeze(-1), 0.0)
    if f32_linear:
        return linear_attention(qf.float(), kf.float(), v.float(), cos_weighting).to(q.dtype)
    return linear_attention(qf, kf, v, cos_weighting)


# --------------------------------------------------------------------------------------------
# randomized attention (randomized_attention.py:24-55)
# --------------------------------------------------------------------------------------------
def ra_probabilities(q, k):
    return torch.softmax(q.shape[-1] ** -0.5 * (q @ k.transpose(-1, -2)), -1)


def ra_core(q, k, v, *, num_samples: int, k_ind=None, noise=None):
    """num_samples 0: mu = q + mean k; -1: mu = q + E_pi[k]; otherwise mu = q + k[k_ind] with ONE key index per query drawn from
    pi = softmax(scale q k^T) (k_ind [B, h, N], explicit here).  The padding mask is ignored, as in the reference."""
    s = q.shape[-1] ** -0.5
    if num_samples == 0:
        mu = q + k.mean(-2, keepdim=True)
    elif num_samples == -1:
        mu = q + ra_probabilities(q, k) @ k
    else:
        mu = q + torch.gather(k, 2, k_ind.unsqueeze(-1).expand(-1, -1, -1, k.shape[-1]))
    w = mu if noise is None else mu + noise
    logits = s * (w @ k.transpose(-1, -2)) - 0.5 * s * (k * k).sum(-1).unsqueeze(-2)
    return torch.softmax(logits, -1) @ v


# --------------------------------------------------------------------------------------------
# ScatterBrain (scatterbrain_attention.py:10-164)
# --------------------------------------------------------------------------------------------
def log_favorp(x, proj):
    d, m = x.shape[-1], proj.shape[1]
    dn = d ** -0.25
    return torch.einsum('bhnd,hjd->bhnj', dn * x, proj) - 0.5 * dn * dn * (x * x).sum(-1, keepdim=True) - 0.5 * math.log(m)


def scatterbrain_core(q, k, v, *, proj, seq_shape, window, ext, pad_mask=None, bias=None):
    """Local window attention + self-normalised random-feature attention over everything OUTSIDE the window, one joint softmax
    (:95-160).  Off-sequence halo slots carry log-feature 0 (the reference pads the partitioned tensor with zeros) and v = 0."""
    B, h, N, d = q.shape
    if pad_mask is None:
        pad_mask = torch.zeros(B, N, dtype=torch.bool)
    pad_mask = pad_mask.bool()
    lq = log_favorp(q, proj)
    lk = log_favorp(k, proj).masked_fill(pad_mask.unsqueeze(1).unsqueeze(-1), float('-inf'))
    if len(seq_shape) == 2:
        qi, ki = group_index_2d(seq_shape[0], seq_shape[1], window, 0), group_index_2d(seq_shape[0], seq_shape[1], window, ext)
    else:
        qi, ki = group_index_1d(N, window, 0, 0), group_index_1d(N, window, ext, ext)
    wq_, wk_, wv_ = take_groups(q, qi), take_groups(k, ki), take_groups(v, ki)
    wlq, wlk = take_groups(lq, qi), take_groups(lk, ki, fill=0.0)
    mx = torch.maximum(lk.amax(-2), wlk.amax((-2, -3))).detach()                  # [B, h, m]
    pk = torch.exp(lk - mx.unsqueeze(-2))                                   # [B, h, N, m]
    wpk = torch.exp(wlk - mx.unsqueeze(-2).unsqueeze(-2))                   # [B, h, G, J, m]
    num = (pk.transpose(-1, -2) @ v).unsqueeze(2) - wpk.transpose(-1, -2) @ wv_          # [B, h, G, m, d]
    den = pk.sum(-2).unsqueeze(2) - wpk.sum(-2)                             # [B, h, G, m]
    kv_stats = num / den.unsqueeze(-1).clamp(min=1e-3)
    glse = torch.logsumexp(lk, -2).unsqueeze(2)                             # [B, h, 1, m]
    llse = torch.logsumexp(wlk, -2)                                         # [B, h, G, m]
    a = torch.maximum(glse, llse)
    nonlocal_lse = a + torch.log(torch.exp(glse - a) - torch.exp(llse - a) + 1e-5)      # attn_utils.py:44-51, mask (1, -1)
    rfa_logits = wlq + nonlocal_lse.unsqueeze(-2)                           # [B, h, G, L, m]
    s = d ** -0.5 * torch.einsum('bhwld,bhwjd->bhwlj', wq_, wk_)
    if bias is not None:
        s = s + bias.unsqueeze(0).unsqueeze(2)
    s = s.masked_fill(take_mask_groups(pad_mask, ki).unsqueeze(1).unsqueeze(-2), float('-inf'))
    J = s.shape[-1]
    p = torch.softmax(torch.cat([s, rfa_logits], -1), -1)
    o_w = torch.einsum('bhwlj,bhwjd->bhwld', p[..., :J], wv_) + torch.einsum('bhwlc,bhwcd->bhwld', p[..., J:], kv_stats)
    o = torch.zeros_like(q)
    o[:, :, qi.reshape(-1)] = o_w.reshape(B, h, -1, d)
    return o


# --------------------------------------------------------------------------------------------
# module-level wrappers: state_dict + config + x -> y
# --------------------------------------------------------------------------------------------
def _proj_from_state(sd, cfg, proj):
    if proj is not None:
        return proj
    for name in ('eval_proj', 'random_proj', 'feature_proj.random_proj'):
        if name in sd:
            return sd[name]
    return None


def performer_forward(sd, cfg, x, pad_mask=None, proj=None, f32_linear=False):
    """KernelizedAttention via MultiheadAttention.forward.  `proj`: the training-mode draw of sample_scheme 'default'."""
    heads = cfg['num_heads']
    B, *shape, C = x.shape
    q, k, v = _split_qkv(x.reshape(B, -1, C), sd, heads)
    method = cfg['proj_method']
    nu = (cfg['approx_attn_dim'] // (C // heads)) // 2
    o = performer_core(q, k, v, method=method, proj=_proj_from_state(sd, cfg, proj), nu=nu,
                       mlp_w=sd.get('feature_proj.phi.0.weight'), mlp_b=sd.get('feature_proj.phi.0.bias'),
                       cos_weighting=cfg.get('cos_weighting', False), pad_mask=pad_mask, f32_linear=f32_linear)
    y = o.transpose(1, 2).reshape((B,) + tuple(shape) + (C,))
    return linear(y, sd['proj.weight'], sd['proj.bias'])


def ra_forward(sd, cfg, x, pad_mask=None, k_ind=None, noise=None):
    heads = cfg['num_heads']
    B, *shape, C = x.shape
    q, k, v = _split_qkv(x.reshape(B, -1, C), sd, heads)
    o = ra_core(q, k, v, num_samples=cfg['num_samples'], k_ind=k_ind, noise=noise)
    y = o.transpose(1, 2).reshape((B,) + tuple(shape) + (C,))
    return linear(y, sd['proj.weight'], sd['proj.bias'])


def scatterbrain_forward(sd, cfg, x, pad_mask=None, proj=None):
    heads = cfg['num_heads']
    B, *shape, C = x.shape
    orig_n = int(math.prod(shape))
    w = cfg['window_size']
    ext = max(1, w // 2) if cfg.get('overlap_window') else 0
    if cfg['attn_2d']:
        xf, mask, seq_shape = x.reshape(B, -1, C), pad_mask, tuple(shape)
    else:
        xf, mask = _pad_tokens(x.reshape(B, -1, C), pad_mask, w)       # x itself is padded (local_attention.py:114-132)
        seq_shape = (xf.shape[1],)
    q, k, v = _split_qkv(xf, sd, heads)
    d = C // heads
    L = w * w if cfg['attn_2d'] else w
    J = (w + 2 * ext) ** 2 if cfg['attn_2d'] else w + 2 * ext
    bias = local_bias_from_state(sd, dict(cfg, ext=ext, use_t5_rpe=False), heads, L, J, d ** -0.5)
    o = scatterbrain_core(q, k, v, proj=_proj_from_state(sd, cfg, proj), seq_shape=seq_shape, window=w, ext=ext, pad_mask=mask,
                          bias=bias)
    y = o.permute(0, 2, 1, 3).reshape((B,) + tuple(seq_shape) + (C,))
    y = linear(y, sd['proj.weight'], sd['proj.bias'])
    return y[..., :orig_n, :]
