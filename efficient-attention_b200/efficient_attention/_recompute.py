"""Backward of the attention cores: `autograd.Function`s whose FORWARD is the libeva_sm100 kernel call.

EVA / causal EVA (`EvaCoreFn`): the BACKWARD is `eva_backward` of libeva_sm100 (csrc/eva_backward.cu: the probabilities are
recomputed tile by tile from q, k, v -- the flash-attention recipe: store inputs and the output, not probabilities); the parameter
gradients of the adaptive Linear / LayerNorm are library reductions over the per-chunk rows the kernel leaves.  `set_backward_impl(
'torch')` (or EVA_SM100_BACKWARD=torch) switches to the older route, autograd through `eva_core_torch`, which the tests keep as a
second opinion.  LARA (`LaraCoreFn`) still differentiates `lara_core_torch`.

`eva_core_torch` / `lara_core_torch` are float32 restatements with differentiable PyTorch ops on [B, N, H, d] views (eva.py:151-227 /
causal_eva.py:676-783 / lara.py:84-251).  They are NEVER used for a forward result, with one labelled exception:
attention-probability dropout (`CausalEVAttention(dropout > 0)` in training mode, causal_eva.py:778) -- the kernels have no dropout,
so that configuration runs `eva_core_torch` for the forward too (`EvaCoreFn` is bypassed; see causal_eva.py here).
"""
import math
import os

import torch
import torch.nn.functional as F

from . import _abi

MASK_VAL = -5.0e4          # eva.py:139, causal_eva.py:488
_BACKWARD_IMPL = os.environ.get('EVA_SM100_BACKWARD', 'cuda')
_LARA_BACKWARD_IMPL = os.environ.get('EVA_SM100_LARA_BACKWARD', 'fused')        # 'fused' | 'explicit' | 'autograd'


def set_backward_impl(name):
    """'cuda' (eva_backward kernels) or 'torch' (autograd through eva_core_torch); returns the previous setting."""
    global _BACKWARD_IMPL
    assert name in ('cuda', 'torch')
    prev, _BACKWARD_IMPL = _BACKWARD_IMPL, name
    return prev


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
    if g is None:
        return y
    # LayerNorm spelled out: the gain / bias gradients then are plain sum reductions.  (F.layer_norm's weight-gradient kernel takes
    # ~0.2 ms per call on the [B * H * C, 64] landmark rows -- 4 ms of a DeiT-small + LARA training step.)
    yf = y.float()
    mu = yf.mean(-1, keepdim=True)
    var = yf.var(-1, unbiased=False, keepdim=True)
    return (yf - mu) * torch.rsqrt(var + eps) * g.float() + beta.float()


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


def _cuda_backward(saved, meta, need, grad_out, packed_out=False):
    """`eva_backward` on the saved (q, k, v, noise, bias, 8 parameters, out); need = needs_input_grad of (bias, 8 parameters).
    Returns (float32 [3, B, N, H, D] = dq | dk | dv -- or, packed_out, q's dtype [B, N, 3, H, D] --, (d bias, 8 parameter gradients))."""
    q, k, v, noise, bias, wq, bq, gq, betq, wk, bk, gk, betk, out = saved[:14]
    stats = tuple(saved[14:]) if len(saved) >= 16 else None            # (k_bar, beta[, row log-sum-exp]) kept by the forward
    geom = _abi.eva_geometry(q, **meta['geometry'])
    ada = _abi.adaptive(wq, bq, gq, betq, wk, bk, gk, betk, mu_coeff=meta['mu_coeff'])
    gqkv, gbias, rows = _abi.eva_backward(q, k, v, geom, ada, out, grad_out, pad_mask=meta['pad_mask'], noise=noise, bias=bias,
                                          want_bias_grad=bias is not None and need[0], stats=stats, packed_out=packed_out)
    B, H, D = q.shape[0], q.shape[2], q.shape[-1]
    # per-chunk rows, grouped per (batch, head): the reductions over ~B*H*C rows run as B*H small GEMMs and one sum
    dyk, dyq, mk, mq, nk, nq, dok, doq = (rows[i].reshape(B * H, -1, D) for i in range(4, 12))
    like = lambda g_, src: g_.to(src.dtype)
    res = [like(gbias, bias) if gbias is not None else None]
    makers = ((wq, lambda: torch.bmm(dyq.transpose(1, 2), mq).sum(0)), (bq, lambda: dyq.sum((0, 1))), (gq, lambda: (doq * nq).sum((0, 1))),
              (betq, lambda: doq.sum((0, 1))), (wk, lambda: torch.bmm(dyk.transpose(1, 2), mk).sum(0)), (bk, lambda: dyk.sum((0, 1))),
              (gk, lambda: (dok * nk).sum((0, 1))), (betk, lambda: dok.sum((0, 1))))
    for i, (src, make) in enumerate(makers):
        res.append(like(make(), src) if (src is not None and need[1 + i]) else None)
    return gqkv, tuple(res)


class EvaCoreFn(torch.autograd.Function):
    """forward: `eva_forward` of libeva_sm100; backward: `eva_backward` (or autograd through `eva_core_torch`, see the module docstring)."""

    @staticmethod
    def forward(ctx, q, k, v, noise, bias, wq, bq, gq, betq, wk, bk, gk, betk, meta):
        geom = _abi.eva_geometry(q, keep_stats=True, **meta['geometry'])
        ada = _abi.adaptive(wq, bq, gq, betq, wk, bk, gk, betk, mu_coeff=meta['mu_coeff'])
        out, stats = _abi.eva_forward(q, k, v, geom, ada, pad_mask=meta['pad_mask'], noise=noise,
                                      bias=None if bias is None else bias.detach(), return_stats=True)
        ctx.meta = meta
        ctx.save_for_backward(q, k, v, noise, bias, wq, bq, gq, betq, wk, bk, gk, betk, out, *stats)
        return out

    @staticmethod
    def _backward_cuda(ctx, grad_out):
        q, k, v = ctx.saved_tensors[:3]
        need = ctx.needs_input_grad
        gqkv, rest = _cuda_backward(ctx.saved_tensors, ctx.meta, need[4:13], grad_out)
        return (gqkv[0].to(q.dtype) if need[0] else None, gqkv[1].to(k.dtype) if need[1] else None,
                gqkv[2].to(v.dtype) if need[2] else None, None) + rest + (None,)

    @staticmethod
    def backward(ctx, grad_out):
        if _BACKWARD_IMPL == 'cuda':
            return EvaCoreFn._backward_cuda(ctx, grad_out)
        saved = ctx.saved_tensors[:13]          # the torch route recomputes everything from the inputs
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


class EvaCorePackedFn(torch.autograd.Function):
    """EvaCoreFn for q, k, v that are the three slices of one packed [B, N, 3, H, d] projection output (abstract_attention.py:72-78):
    the gradient goes back as ONE tensor in that layout, written by the backward kernels themselves (`grad_qkv_io` of eva_backward),
    instead of three casts followed by autograd's three zero-filled `select` gradients and their sums."""

    @staticmethod
    def forward(ctx, packed, noise, bias, wq, bq, gq, betq, wk, bk, gk, betk, meta):
        q, k, v = packed[:, :, 0], packed[:, :, 1], packed[:, :, 2]
        geom = _abi.eva_geometry(q, keep_stats=True, **meta['geometry'])
        ada = _abi.adaptive(wq, bq, gq, betq, wk, bk, gk, betk, mu_coeff=meta['mu_coeff'])
        out, stats = _abi.eva_forward(q, k, v, geom, ada, pad_mask=meta['pad_mask'], noise=noise,
                                      bias=None if bias is None else bias.detach(), return_stats=True)
        ctx.meta = meta
        ctx.save_for_backward(packed, noise, bias, wq, bq, gq, betq, wk, bk, gk, betk, out, *stats)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        packed = ctx.saved_tensors[0]
        need = ctx.needs_input_grad
        saved = (packed[:, :, 0], packed[:, :, 1], packed[:, :, 2]) + tuple(ctx.saved_tensors[1:])
        gp, rest = _cuda_backward(saved, ctx.meta, need[2:11], grad_out, packed_out=True)
        return (gp if need[0] else None, None) + rest + (None,)


def eva_core(q, k, v, *, geometry, mu_coeff, params, pad_mask=None, noise=None, bias=None, packed=None):
    """Kernel forward + backward.  `params` = (wq, bq, gq, betq, wk, bk, gk, betk), entries may be None.  `packed`: the contiguous
    [B, N, 3, H, d] tensor q, k, v are slices of, when there is one."""
    meta = dict(geometry=geometry, mu_coeff=mu_coeff, pad_mask=pad_mask)
    if packed is not None and _BACKWARD_IMPL == 'cuda' and packed.dim() == 5 and packed.shape[2] == 3 and packed.is_contiguous():
        return EvaCorePackedFn.apply(packed, noise, bias, *params, meta)
    return EvaCoreFn.apply(q, k, v, noise, bias, *params, meta)


class WindowCoreFn(torch.autograd.Function):
    """Chunk-less window attention (local_attention.py:134-182; the dense softmax baseline abstract_attention.py:115-133 is the
    one-window case): forward `eva_window_attention`, backward `eva_backward` of libeva_sm100."""

    @staticmethod
    def forward(ctx, q, k, v, bias, meta):
        geom = _abi.eva_geometry(q, **meta['geometry'])
        out, lse = _abi.eva_window_attention(q, k, v, geom, pad_mask=meta['pad_mask'], bias=None if bias is None else bias.detach(),
                                             return_lse=True)
        ctx.meta = meta
        ctx.has_lse = lse is not None
        ctx.save_for_backward(q, k, v, bias, out, *(() if lse is None else (lse,)))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        q, k, v, bias, out = ctx.saved_tensors[:5]
        lse = ctx.saved_tensors[5] if ctx.has_lse else None
        need = ctx.needs_input_grad
        geom = _abi.eva_geometry(q, **ctx.meta['geometry'])
        gqkv, gbias, _ = _abi.eva_backward(q, k, v, geom, None, out, grad_out, pad_mask=ctx.meta['pad_mask'], bias=bias,
                                           want_bias_grad=bias is not None and need[3], lse=lse)
        return (gqkv[0].to(q.dtype) if need[0] else None, gqkv[1].to(k.dtype) if need[1] else None,
                gqkv[2].to(v.dtype) if need[2] else None, None if gbias is None else gbias.to(bias.dtype), None)


def window_core(q, k, v, *, geometry, pad_mask=None, bias=None):
    return WindowCoreFn.apply(q, k, v, bias, dict(geometry=geometry, pad_mask=pad_mask))


def needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def num_chunks_of(seq_shape, chunk, chunk_ext=0):
    if len(seq_shape) == 2:
        return (seq_shape[0] // chunk) * (seq_shape[1] // chunk)
    return seq_shape[0] // chunk


# ------------------------------------------------------------------------------------------------
# LARA (lara.py:84-251)
# ------------------------------------------------------------------------------------------------
def _prm(data, omega):
    """d^-1/2 <omega_c, x_n> - d^-1/2 |x_n|^2 / 2 -> [..., c, n]   (attn_utils.py:324-336, 347)."""
    s = data.shape[-1] ** -0.5
    return s * (omega @ data.transpose(-1, -2)) - 0.5 * s * (data * data).sum(-1).unsqueeze(-2)


def _segment_means(t, n_lm):
    """1-D landmarks: means over n_lm consecutive segments, the first ones one token shorter when N % n_lm != 0 (lara.py:104-124)."""
    B, H, N, d = t.shape
    if N <= n_lm:
        return t
    seg = N // n_lm
    if N % n_lm == 0:
        return t.reshape(B, H, n_lm, seg, d).mean(-2)
    n_short = (seg + 1) * n_lm - N
    a = t[:, :, :n_short * seg].reshape(B, H, n_short, seg, d).mean(-2)
    b = t[:, :, n_short * seg:].reshape(B, H, n_lm - n_short, seg + 1, d).mean(-2)
    return torch.cat([a, b], -2)


def _lara_stage1(q, k, v, *, seq_shape, landmarks, per_token_proj, mixed, sample_mode, zero_padded, wq, bq, gq, betq, wk, bk, gk, betk,
                 dense, ln_eps, pad_mask, noise, keep_dtype):
    """Everything of LARA that lives on [B, H, C, d] tensors (lara.py:84-198): landmarks q_bar / k_bar (pooled 2-D 'light' / 'dense', or
    1-D segment means with optional per-token Linear + LayerNorm), optional landmark mixing, mu, the proposal samples omega.
    Returns (qh, kh, vh [B, H, N, d], q_bar, mu, omega, rep)."""
    B, N, H, d = q.shape
    scale = d ** -0.5
    # keep_dtype: leave 16-bit q, k, v as they are -- the caller runs this under torch.autocast, i.e. with the numerics the reference
    # itself trains with (matmuls in the 16-bit format, softmax / logsumexp / LayerNorm / sums in float32)
    qh, kh, vh = ((t if keep_dtype else t.float()).permute(0, 2, 1, 3) for t in (q, k, v))           # [B, H, N, d]
    if zero_padded and pad_mask is not None:
        keep = (~pad_mask.to(torch.bool)).to(qh.dtype).view(B, 1, N, 1)
        qh, kh, vh = qh * keep, kh * keep, vh * keep
    two_d = len(seq_shape) == 2
    if two_d:
        gh, gw = seq_shape
        side = int(math.isqrt(landmarks))

        def pool(t):                                                            # [B, H, N, d] -> [B, H, side^2, d]
            grid = t.reshape(B * H, gh, gw, d).permute(0, 3, 1, 2)
            return F.adaptive_avg_pool2d(grid, side).flatten(2).transpose(1, 2).reshape(B, H, side * side, d)
        q_bar, k_bar = pool(qh), pool(kh)
        if wq is not None and dense:                                            # Linear / LayerNorm over all channels (head-major)
            merge = lambda t: t.permute(0, 2, 1, 3).reshape(B, -1, H * d)
            split = lambda t: t.reshape(B, -1, H, d).permute(0, 2, 1, 3)
            q_bar = split(_linear_ln(merge(q_bar), wq, bq, gq, betq, ln_eps))
            k_bar = split(_linear_ln(merge(k_bar), wk, bk, gk, betk, ln_eps))
        elif wq is not None:
            q_bar, k_bar = _linear_ln(q_bar, wq, bq, gq, betq, ln_eps), _linear_ln(k_bar, wk, bk, gk, betk, ln_eps)
        if mixed:
            lg = scale * (k_bar @ k_bar.transpose(-1, -2))
            if mixed == 2:
                lg = lg + torch.log(torch.linalg.vector_norm(pool(vh), ord=2, dim=-1) + 1e-4).unsqueeze(-2)
            k_bar = torch.softmax(lg, -1) @ k_bar
    else:
        q2, k2 = qh, kh
        if per_token_proj:
            q2, k2 = _linear_ln(qh, wq, bq, gq, betq, ln_eps), _linear_ln(kh, wk, bk, gk, betk, ln_eps)
        q_bar, k_bar = _segment_means(q2, landmarks), _segment_means(k2, landmarks)
    mu = q_bar + k_bar
    rep = 1
    if noise is None:
        omega = mu
    elif sample_mode == _abi.LARA_SAMPLE_ANTITHETIC:
        omega, rep = torch.cat([mu + noise.float(), mu - noise.float()], -2), 2
    elif sample_mode == _abi.LARA_SAMPLE_MULTI:
        omega, rep = mu.repeat(1, 1, 2, 1) + noise.float(), 2
    else:
        omega = mu + noise.float()
    return qh, kh, vh, q_bar, mu, omega, rep


def lara_core_torch(q, k, v, *, seq_shape, landmarks, per_token_proj, mixed, mis_type, sample_mode, zero_padded, alpha_coeff,
                    wq, bq, gq, betq, wk, bk, gk, betk, dense=False, ln_eps=1e-5, pad_mask=None, noise=None, keep_dtype=False):
    """q, k, v [B, N, H, d] -> [B, N, H * d] float32: `_lara_stage1`, then the self-normalised importance-sampling estimator with the
    three MIS variants (lara.py:201-246).  sample_mode: 0 one sample per landmark, 1 antithetic (noise [.., C, d] -> [mu + e ; mu - e]),
    2 multi (noise [.., 2C, d])."""
    B, N, H, d = q.shape
    scale = d ** -0.5
    qh, kh, vh, q_bar, mu, omega, rep = _lara_stage1(
        q, k, v, seq_shape=seq_shape, landmarks=landmarks, per_token_proj=per_token_proj, mixed=mixed, sample_mode=sample_mode,
        zero_padded=zero_padded, wq=wq, bq=bq, gq=gq, betq=betq, wk=wk, bk=bk, gk=gk, betk=betk, dense=dense, ln_eps=ln_eps,
        pad_mask=pad_mask, noise=noise, keep_dtype=keep_dtype)
    A = _prm(qh, omega)                                                         # [B, H, S, N]
    Bk = _prm(kh, omega)
    if pad_mask is not None:
        Bk = Bk.masked_fill(pad_mask.to(torch.bool).view(B, 1, 1, N), float('-inf'))
    kv = torch.softmax(Bk, -1) @ vh
    if mis_type == 'mis-opt':
        t = torch.softmax(scale * (q_bar @ qh.transpose(-1, -2)), -1).repeat(1, 1, rep, 1)
        Lm = _prm(mu.repeat(1, 1, rep, 1), omega)                               # [B, H, S(omega), S(mu)]
        lp = torch.diagonal(Lm, dim1=-1, dim2=-2).unsqueeze(-1)
        bh = torch.exp(lp - torch.logsumexp(Lm, -1, keepdim=True))
        log_alpha = torch.log((bh + alpha_coeff * (t - t.mean(-2, keepdim=True))).clamp(min=1e-8))
    elif mis_type == 'mis-bh':
        log_alpha = 0.0
        lp = torch.logsumexp(_prm(mu, omega), -1, keepdim=True)
    else:                                                                       # 'mis-biased'
        log_alpha = (scale * (mu @ qh.transpose(-1, -2))).repeat(1, 1, rep, 1)
        lp = torch.logsumexp(_prm(mu, omega), -1, keepdim=True)
    logw = log_alpha + A + torch.logsumexp(Bk, -1, keepdim=True) - lp
    out = torch.softmax(logw, -2).transpose(-1, -2) @ kv                        # [B, H, N, d]
    return out.permute(0, 2, 1, 3).reshape(B, N, H * d)


def _lara_stage2_backward(qf, kf, vf, gf, q_bar, omega, lp, bh, coeff):
    """Explicit gradient of the mis-opt estimator (lara.py:201-246, one sample per landmark, no padding mask) with respect to
    q, k, v [B, H, N, d] and to the small inputs q_bar, omega [B, H, C, d], lp, bh [B, H, C, 1]; everything float32.  The forward
    quantities are recomputed here -- nothing of size [C, N] is kept between the forward kernel and this call.
        A = s (omega q^T) - s |q|^2 / 2,   T = s q_bar q^T,  t = softmax_n T,   Bk = s (omega k^T) - s |k|^2 / 2,  Pk = softmax_m Bk,
        kv = Pk v,  alpha = bh + coeff (t - mean_c t),  logw = log max(alpha, 1e-8) + A + lse_m Bk - lp,  W = softmax_c logw,  O = W^T kv"""
    d = qf.shape[-1]
    s = d ** -0.5
    q2 = (0.5 * s) * (qf * qf).sum(-1).unsqueeze(-2)                            # [B, H, 1, N]
    k2 = (0.5 * s) * (kf * kf).sum(-1).unsqueeze(-2)
    A = s * (omega @ qf.transpose(-1, -2)) - q2                                # [B, H, C, N]
    T = s * (q_bar @ qf.transpose(-1, -2))
    Bk = s * (omega @ kf.transpose(-1, -2)) - k2
    lseB = torch.logsumexp(Bk, -1, keepdim=True)
    Pk = torch.exp(Bk - lseB)
    t = torch.softmax(T, -1)
    kv = Pk @ vf                                                                # [B, H, C, d]
    alpha = bh + coeff * (t - t.mean(-2, keepdim=True))
    W = torch.softmax(torch.log(alpha.clamp(min=1e-8)) + A + lseB - lp, -2)     # [B, H, C, N]
    dW = kv @ gf.transpose(-1, -2)
    dlw = W * (dW - (W * dW).sum(-2, keepdim=True))
    dkv = W @ gf
    dlseB = dlw.sum(-1, keepdim=True)
    dalpha = torch.where(alpha > 1e-8, dlw / alpha, torch.zeros_like(dlw))
    dbh = dalpha.sum(-1, keepdim=True)
    dt = coeff * (dalpha - dalpha.mean(-2, keepdim=True))
    dT = t * (dt - (t * dt).sum(-1, keepdim=True))
    dBk = Pk * (dkv @ vf.transpose(-1, -2) - (dkv * kv).sum(-1, keepdim=True) + dlseB)
    domega = s * (dlw @ qf + dBk @ kf)
    dqbar = s * (dT @ qf)
    dq = s * (dlw.transpose(-1, -2) @ omega + dT.transpose(-1, -2) @ q_bar - dlw.sum(-2).unsqueeze(-1) * qf)
    dk = s * (dBk.transpose(-1, -2) @ omega - dBk.sum(-2).unsqueeze(-1) * kf)
    dv = Pk.transpose(-1, -2) @ dkv
    return dq, dk, dv, dqbar, domega, -dlseB, dbh


def _lara_stage2_backward_fused(qc, kc, vc, gc, q_bar, omega, lp, bh, coeff):
    """`_lara_stage2_backward` with the row / column softmax algebra in three fused kernels (`lara_backward_step` of libeva_sm100)
    and the GEMMs as batched library calls in the activation format.  qc, kc, vc, gc: [BH, N, d] contiguous (16-bit or float32);
    q_bar, omega [BH, C, d], lp, bh [BH, C] float32.  Returns float32 (dq, dk, dv [BH, N, d], d q_bar, d omega [BH, C, d], d lp, d bh
    [BH, C])."""
    BH, N, d = qc.shape
    C = omega.shape[1]
    s = d ** -0.5
    dt_ = qc.dtype
    f32 = dict(dtype=torch.float32, device=qc.device)
    OQ = torch.cat([omega, q_bar], 1).to(dt_)                                   # [BH, 2C, d]
    om = OQ[:, :C]
    X = OQ @ qc.transpose(1, 2)                                                 # [BH, 2C, N]: omega q^T | q_bar q^T
    Bm = om @ kc.transpose(1, 2)                                                # [BH, C, N]:  omega k^T
    q2s = (0.5 * s) * (qc.float() ** 2).sum(-1)
    k2s = (0.5 * s) * (kc.float() ** 2).sum(-1)
    lseB, lseT = torch.empty(BH, C, **f32), torch.empty(BH, C, **f32)
    kw = dict(items=BH, landmarks=C, tokens=N, scale=s)
    _abi.lara_backward_step(0, X, Bm, v0=k2s, o0=lseB, o1=lseT, **kw)           # X[:, C:] = t, Bm = Pk
    kv = Bm @ vc                                                                # [BH, C, d]
    dW = kv @ gc.transpose(1, 2)                                                # [BH, C, N]
    M2 = torch.empty(BH, 2 * C, N, dtype=dt_, device=qc.device)
    sums = torch.zeros(3, BH, C, **f32)                                         # d lse_B | d bh | R
    _abi.lara_backward_step(1, X, None, dW=dW, M2=M2, v0=q2s, v1=bh.contiguous(), v2=lp.contiguous(), v3=lseB, o0=sums[0], o1=sums[1],
                            o2=sums[2], alpha_coeff=coeff, **kw)                # X[:, :C] = W, M2 = [dlw ; dt]
    dlseB, dbh, R = sums[0], sums[1], sums[2]
    _abi.lara_backward_step(2, M2[:, C:], X[:, C:], v0=R, x_item_stride=2 * C * N, y_item_stride=2 * C * N, **kw)       # dT = t (dt - R)
    dkv = X[:, :C] @ gc                                                         # [BH, C, d]
    E = (dkv.float() * kv.float()).sum(-1)
    dBk = dkv @ vc.transpose(1, 2)                                              # dPk, then dBk in place
    _abi.lara_backward_step(2, dBk, Bm, v0=E, v1=dlseB, x_item_stride=C * N, y_item_stride=C * N, **kw)
    dOQ = s * (M2 @ qc).float()                                                 # [BH, 2C, d]: d omega (A part) | d q_bar
    domega = dOQ[:, :C] + s * (dBk @ kc).float()
    dq = s * (M2.transpose(1, 2) @ OQ).float()                                  # the - sum_c dlw q_n term vanishes: columns of dlw sum to 0
    dk = s * ((dBk.transpose(1, 2) @ om).float() - dBk.float().sum(1).unsqueeze(-1) * kc.float())
    dv = (Bm.transpose(1, 2) @ dkv).float()
    return dq, dk, dv, dOQ[:, C:], domega, -dlseB, dbh


class LaraCoreFn(torch.autograd.Function):
    """forward: `lara_forward` of libeva_sm100 (landmarks given by the caller when `dense`); backward: autograd through
    `lara_core_torch` on the saved inputs."""

    @staticmethod
    def forward(ctx, q, k, v, noise, given, wq, bq, gq, betq, wk, bk, gk, betk, meta):
        proj = _abi.adaptive(*([None] * 8), mu_coeff=1.0) if (meta['dense'] or wq is None) else \
            _abi.adaptive(wq, bq, gq, betq, wk, bk, gk, betk, mu_coeff=1.0, ln_eps=meta['ln_eps'])
        out = _abi.lara_forward(q, k, v, proj=proj, pad_mask=meta['pad_mask'], noise=noise,
                                given_landmarks=None if given is None else given.detach(), **meta['kernel'])
        ctx.meta = meta
        ctx.save_for_backward(q, k, v, noise, wq, bq, gq, betq, wk, bk, gk, betk)
        return out

    @staticmethod
    def _backward_explicit(saved, meta, need, grad_out):
        """mis-opt, one sample per landmark, no padding mask: the [C, N]-sized part of the estimator is differentiated by explicit
        formulas (`_lara_stage2_backward`: batched GEMMs + row / column softmax algebra, float32), only the landmark stage
        (`_lara_stage1`, tensors of size [C, d] plus the pooling) goes through autograd."""
        kk = meta['kernel']
        # 16-bit activations with the fused kernels: the landmark stage runs under autocast on the 16-bit q, k, v (no float32 copies of
        # the activations); its outputs (LayerNorm / softmax results) are float32 either way
        half = _LARA_BACKWARD_IMPL == 'fused' and saved[0].dtype in (torch.float16, torch.bfloat16)
        with torch.enable_grad(), torch.autocast('cuda', dtype=saved[0].dtype if half else torch.float16, enabled=half):
            ins = [None if t is None else t.detach().requires_grad_(n and t.is_floating_point()) for t, n in zip(saved, need)]
            q, k, v, noise, wq, bq, gq, betq, wk, bk, gk, betk = ins
            qh, kh, vh, q_bar, mu, omega, _ = _lara_stage1(
                q, k, v, seq_shape=kk['seq_shape'], landmarks=kk['landmarks'], per_token_proj=kk['per_token_proj'], mixed=kk['mixed'],
                sample_mode=kk['sample_mode'], zero_padded=False, wq=wq, bq=bq, gq=gq, betq=betq, wk=wk, bk=bk, gk=gk, betk=betk,
                dense=meta['dense'], ln_eps=meta['ln_eps'], pad_mask=None, noise=noise, keep_dtype=half)
            q_bar, mu, omega = q_bar.float(), mu.float(), omega.float()
            Lm = _prm(mu, omega)
            lp = torch.diagonal(Lm, dim1=-1, dim2=-2).unsqueeze(-1)
            bh = torch.exp(lp - torch.logsumexp(Lm, -1, keepdim=True))
        B, N, H, d = saved[0].shape
        with torch.no_grad():
            if _LARA_BACKWARD_IMPL == 'fused':
                C = omega.shape[2]
                flat = lambda t: t.detach().permute(0, 2, 1, 3).reshape(B * H, N, d).contiguous()          # [B, N, H, d] -> [BH, N, d]
                gq = grad_out.reshape(B, N, H, d).to(saved[0].dtype)
                r = _lara_stage2_backward_fused(flat(saved[0]), flat(saved[1]), flat(saved[2]), flat(gq), q_bar.detach().float().reshape(B * H, C, d),
                                                omega.detach().float().reshape(B * H, C, d), lp.detach().float().reshape(B * H, C),
                                                bh.detach().float().reshape(B * H, C), kk['alpha_coeff'])
                dq2, dk2, dv2 = (t.view(B, H, N, d) for t in r[:3])
                dqbar, domega = r[3].view(B, H, C, d), r[4].view(B, H, C, d)
                dlp, dbh = r[5].view(B, H, C, 1), r[6].view(B, H, C, 1)
            else:
                gf = grad_out.reshape(B, N, H, d).permute(0, 2, 1, 3).float()
                dq2, dk2, dv2, dqbar, domega, dlp, dbh = _lara_stage2_backward(
                    qh.detach().contiguous(), kh.detach().contiguous(), vh.detach().contiguous(), gf.contiguous(), q_bar.detach(),
                    omega.detach(), lp.detach(), bh.detach(), kk['alpha_coeff'])
        wanted = [t for t in ins if t is not None and t.requires_grad]
        g1 = torch.autograd.grad([q_bar, omega, lp, bh], wanted, [dqbar.to(q_bar.dtype), domega.to(omega.dtype), dlp.to(lp.dtype),
                                                                   dbh.to(bh.dtype)], allow_unused=True)
        direct = {id(q): dq2, id(k): dk2, id(v): dv2}
        it = iter(g1)
        res = []
        for t, src in zip(ins, saved):
            if t is not None and t.requires_grad:
                gr = next(it)
                extra = direct.get(id(t))
                if extra is not None:
                    extra = extra.permute(0, 2, 1, 3)                      # [B, H, N, d] -> [B, N, H, d]
                    gr = extra if gr is None else gr + extra
                res.append(None if gr is None else gr.to(src.dtype))
            else:
                res.append(None)
        return tuple(res[:4]) + (None,) + tuple(res[4:]) + (None,)

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        meta = ctx.meta
        need = list(ctx.needs_input_grad[:4]) + list(ctx.needs_input_grad[5:13])
        kk0 = meta['kernel']
        # float32 activations: explicit float32 formulas for the [C, N]-sized part (exact, nothing of that size kept by autograd);
        # 16-bit activations: autograd under autocast below is as fast (tensor-core GEMMs) and is what the reference trains with
        # (measured, DeiT-small-p16 + LARA step at batch 128, fp16: 66.3 ms autograd / autocast, 68.0 ms explicit float32)
        if (_LARA_BACKWARD_IMPL in ('explicit', 'fused') and (_LARA_BACKWARD_IMPL == 'fused' or saved[0].dtype == torch.float32) and
                kk0['mis_type'] == 'mis-opt' and meta['pad_mask'] is None and
                (saved[3] is None or kk0['sample_mode'] == _abi.LARA_SAMPLE_SINGLE)):
            return LaraCoreFn._backward_explicit(saved, meta, need, grad_out)
        half = saved[0].dtype in (torch.float16, torch.bfloat16)       # 16-bit activations: differentiate under autocast, as the reference trains
        with torch.enable_grad(), torch.autocast('cuda', dtype=saved[0].dtype if half else torch.float16, enabled=half):
            ins = [None if t is None else t.detach().requires_grad_(n and t.is_floating_point()) for t, n in zip(saved, need)]
            q, k, v, noise, wq, bq, gq, betq, wk, bk, gk, betk = ins
            kk = meta['kernel']
            out = lara_core_torch(q, k, v, seq_shape=kk['seq_shape'], landmarks=kk['landmarks'], per_token_proj=kk['per_token_proj'],
                                  mixed=kk['mixed'], mis_type=kk['mis_type'], sample_mode=kk['sample_mode'],
                                  zero_padded=kk['zero_padded'], alpha_coeff=kk['alpha_coeff'], wq=wq, bq=bq, gq=gq, betq=betq,
                                  wk=wk, bk=bk, gk=gk, betk=betk, dense=meta['dense'], ln_eps=meta['ln_eps'],
                                  pad_mask=meta['pad_mask'], noise=noise, keep_dtype=half)
            wanted = [t for t in ins if t is not None and t.requires_grad]
            grads = torch.autograd.grad(out, wanted, grad_out.to(out.dtype).reshape(out.shape), allow_unused=True)
        it = iter(grads)
        res = []
        for t, src in zip(ins, saved):
            if t is not None and t.requires_grad:
                gr = next(it)
                res.append(None if gr is None else gr.to(src.dtype))
            else:
                res.append(None)
        # inputs were (q, k, v, noise, given, 8 params, meta): `given` is a function of q, k and the parameters -- its gradient is
        # already accounted for by recomputing the landmarks from q, k inside lara_core_torch
        return tuple(res[:4]) + (None,) + tuple(res[4:]) + (None,)


def lara_core(q, k, v, *, kernel_args, params, dense, ln_eps, pad_mask=None, noise=None, given=None):
    meta = dict(kernel=kernel_args, dense=dense, ln_eps=ln_eps, pad_mask=pad_mask)
    return LaraCoreFn.apply(q, k, v, noise, given, *params, meta)
