"""Backward of the EVA attention cores: `autograd.Function`s whose FORWARD is the libeva_sm100 kernel call and
whose BACKWARD recomputes the core in float32 with differentiable PyTorch ops on the device and lets autograd differentiate
that recomputation (the flash-attention recipe -- store inputs, not probabilities -- with library ops standing in for
hand-written backward kernels, which are the next step).

`eva_core_torch` is that recomputation: the maths of eva.py:151-227 / causal_eva.py:676-783 on [B, N, H, d] views, written with
gather-index tables.  It is NEVER used for a forward result, with one labelled exception: attention-probability dropout
(`CausalEVAttention(dropout > 0)` in training mode, causal_eva.py:778) -- the kernels have no dropout, so that configuration
runs the recomputation for the forward too (`EvaCoreFn` is bypassed; see causal_eva.py here).
"""
import math

import torch
import torch.nn.functional as F

from . import _abi

MASK_VAL = -5.0e4          # eva.py:139, causal_eva.py:488


def _groups_1d(n, size, left, right, device):
    """[n // size, left + size + right] token ids, -1 off the sequence (attn_utils.py:155-166, causal_eva.py:102-113)."""
    g = torch.arange(n // size, device=device).unsqueeze(1) * size - left
    idx = g + torch.arange(left + size + right, device=device).unsqueeze(0)
    return torch.where((idx < 0) | (idx >= n), torch.full_like(idx, -1), idx)


def _groups_2d(gh, gw, size, ext, device):
    """[(gh // size) * (gw // size), (size + 2 ext)^2]: groups row-major, slots row-major (attn_utils.py:172-210)."""
    ny, nx, t = gh // size, gw // size, size + 2 * ext
    ar = lambda n: torch.arange(n, device=device)
    yy = (ar(ny) * size - ext).view(ny, 1, 1, 1) + ar(t).view(1, 1, t, 1)
    xx = (ar(nx) * size - ext).view(1, nx, 1, 1) + ar(t).view(1, 1, 1, t)
    ok = (yy >= 0) & (yy < gh) & (xx >= 0) & (xx < gw)
    return torch.where(ok, yy * gw + xx, torch.full_like(yy * gw + xx, -1)).reshape(ny * nx, t * t)


def _take(t, idx):
    """t [B, N, H, d] -> [B, H, G, S, d]; off-sequence slots are zero."""
    B, N, H, d = t.shape
    G, S = idx.shape
    flat = idx.clamp(min=0).reshape(-1)
    out = t.index_select(1, flat).view(B, G, S, H, d).permute(0, 3, 1, 2, 4)
    return out * (idx >= 0).to(t.dtype).view(1, 1, G, S, 1)


def _take_mask(mask, idx):
    """mask [B, N] bool -> [B, G, S] bool, True where padded or off the sequence."""
    B = mask.shape[0]
    G, S = idx.shape
    return mask.index_select(1, idx.clamp(min=0).reshape(-1)).view(B, G, S) | (idx < 0).view(1, G, S)


def _linear_ln(x, w, b, g, beta, eps):
    y = F.linear(x, w, b)
    return y if g is None else F.layer_norm(y, (y.shape[-1],), g, beta, eps)


def eva_core_torch(q, k, v, *, seq_shape, window, ext, chunk, chunk_ext, wq, bq, gq, betq, wk, bk, gk, betk, mu_coeff,
                   ln_eps=1e-5, pad_mask=None, noise=None, bias=None, causal=False, left_only=False, mask_queries=False,
                   p_drop=0.0, drop_mask=None):
    """q, k, v [B, N, H, d] (any float dtype; computed in float32) -> [B, N, H * d] float32.  bias [1 or H, L, J] already scaled.
    wq None: adaptive_proj == 'none' (mu = 0).  p_drop / drop_mask: dropout on the joint probabilities (causal_eva.py:778);
    drop_mask (bool, True = keep, [B, H, W, L, J + C]) makes the draw explicit."""
    B, N, H, d = q.shape
    dev = q.device
    q, k, v = q.float(), k.float(), v.float()
    scale = d ** -0.5
    if len(seq_shape) == 2:
        gh, gw = seq_shape
        qi, ki, ci = _groups_2d(gh, gw, window, 0, dev), _groups_2d(gh, gw, window, ext, dev), _groups_2d(gh, gw, chunk, chunk_ext, dev)
    else:
        qi = _groups_1d(N, window, 0, 0, dev)
        ki = _groups_1d(N, window, ext, 0 if left_only else ext, dev)
        ci = _groups_1d(N, chunk, chunk_ext, 0 if left_only else chunk_ext, dev)
    W, L = qi.shape
    J, C = ki.shape[1], ci.shape[0]
    mask = torch.zeros(B, N, dtype=torch.bool, device=dev) if pad_mask is None else pad_mask.to(torch.bool)
    # ---- chunk statistics (eva.py:155-196): masked / off-sequence tokens are zero and still count in the mean ----
    cm = _take_mask(mask, ci)                                           # [B, C, Jc]
    keep = (~cm).to(torch.float32).view(B, 1, C, -1, 1)
    ck, cv = _take(k, ci) * keep, _take(v, ci) * keep                   # [B, H, C, Jc, d]
    k_bar = _linear_ln(ck.mean(-2), wk, bk, gk, betk, ln_eps)           # [B, H, C, d]
    if wq is not None:
        mu = mu_coeff * (_linear_ln((_take(q, ci) * keep).mean(-2), wq, bq, gq, betq, ln_eps) + k_bar)
    else:
        mu = torch.zeros_like(k_bar)
    omega = mu if noise is None else mu + noise.float()
    lg = scale * (torch.einsum('bhcd,bhcjd->bhcj', omega, ck) - 0.5 * (ck * ck).sum(-1))
    lg = lg.masked_fill(cm.unsqueeze(1), MASK_VAL)
    beta = torch.einsum('bhcj,bhcjd->bhcd', torch.softmax(lg, -1), cv)
    # ---- local + chunk logits under one softmax (eva.py:200-227) ----
    wq_, wk_, wv_ = _take(q, qi), _take(k, ki), _take(v, ki)
    r = scale * torch.einsum('bhwld,bhcd->bhwlc', wq_, k_bar)
    s = scale * torch.einsum('bhwld,bhwjd->bhwlj', wq_, wk_)
    if bias is not None:
        s = s + bias.float().view(1, bias.shape[0], 1, L, J)
    km = _take_mask(mask, ki).view(B, 1, W, 1, J)
    if mask_queries:
        km = km | _take_mask(mask, qi).view(B, 1, W, L, 1)
    s = s.masked_fill(km, MASK_VAL)
    if causal:
        s = s.masked_fill(torch.ones(L, J, dtype=torch.bool, device=dev).triu(1 + ext), MASK_VAL)
        hide = torch.arange(C, device=dev).view(1, 1, C) >= (qi // chunk).unsqueeze(-1)      # chunk c visible only if c < chunk(query)
        r = r.masked_fill(hide, MASK_VAL)
    p = torch.softmax(torch.cat([s, r], -1), -1)
    if drop_mask is not None:
        p = p * drop_mask.to(p.dtype) / (1.0 - p_drop)
    elif p_drop > 0.0:
        p = F.dropout(p, p_drop, training=True)
    o = torch.einsum('bhwlj,bhwjd->bhwld', p[..., :J], wv_) + torch.einsum('bhwlc,bhcd->bhwld', p[..., J:], beta)
    out = torch.zeros(B, N, H, d, dtype=torch.float32, device=dev)
    out = out.index_copy(1, qi.reshape(-1), o.permute(0, 2, 3, 1, 4).reshape(B, W * L, H, d))
    return out.reshape(B, N, H * d)


class EvaCoreFn(torch.autograd.Function):
    """forward: `eva_forward` of libeva_sm100; backward: autograd through `eva_core_torch` on the saved inputs."""

    @staticmethod
    def forward(ctx, q, k, v, noise, bias, wq, bq, gq, betq, wk, bk, gk, betk, meta):
        geom = _abi.eva_geometry(q, **meta['geometry'])
        ada = _abi.adaptive(wq, bq, gq, betq, wk, bk, gk, betk, mu_coeff=meta['mu_coeff'])
        out = _abi.eva_forward(q, k, v, geom, ada, pad_mask=meta['pad_mask'], noise=noise,
                               bias=None if bias is None else bias.detach())
        ctx.meta = meta
        ctx.save_for_backward(q, k, v, noise, bias, wq, bq, gq, betq, wk, bk, gk, betk)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        meta = ctx.meta
        need = ctx.needs_input_grad[:13]
        with torch.enable_grad():
            ins = [None if t is None else t.detach().requires_grad_(n and t.is_floating_point()) for t, n in zip(saved, need)]
            q, k, v, noise, bias, wq, bq, gq, betq, wk, bk, gk, betk = ins
            g = meta['geometry']
            out = eva_core_torch(q, k, v, seq_shape=g['seq_shape'], window=g['window'], ext=g['ext'], chunk=g['chunk'],
                                 chunk_ext=g['chunk_ext'], wq=wq, bq=bq, gq=gq, betq=betq, wk=wk, bk=bk, gk=gk, betk=betk,
                                 mu_coeff=meta['mu_coeff'], pad_mask=meta['pad_mask'], noise=noise, bias=bias,
                                 causal=g.get('causal', False), left_only=g.get('halo_left_only', False),
                                 mask_queries=g.get('mask_queries', False))
            wanted = [t for t in ins if t is not None and t.requires_grad]
            grads = torch.autograd.grad(out, wanted, grad_out.float().reshape(out.shape), allow_unused=True)
        it = iter(grads)
        result = []
        for t, src in zip(ins, saved):
            if t is not None and t.requires_grad:
                gr = next(it)
                result.append(None if gr is None else gr.to(src.dtype))
            else:
                result.append(None)
        return tuple(result) + (None,)


def eva_core(q, k, v, *, geometry, mu_coeff, params, pad_mask=None, noise=None, bias=None):
    """Kernel forward + recomputation backward.  `params` = (wq, bq, gq, betq, wk, bk, gk, betk), entries may be None."""
    meta = dict(geometry=geometry, mu_coeff=mu_coeff, pad_mask=pad_mask)
    return EvaCoreFn.apply(q, k, v, noise, bias, *params, meta)


def needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def num_chunks_of(seq_shape, chunk, chunk_ext=0):
    if len(seq_shape) == 2:
        return (seq_shape[0] // chunk) * (seq_shape[1] // chunk)
    return seq_shape[0] // chunk
