"""CPU oracle for the EVA / LARA / causal-EVA attention forward path.

*** TEST INFRASTRUCTURE ONLY. ***  Nothing under ``oracle/`` is imported by the product
package (``efficient-attention_b200/``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker / the CPU baseline, never as the thing shipped.

What this is: an independent restatement (plain torch CPU tensor arithmetic, any float dtype,
normally float64) of the algorithm the reference implements in

    efficient-attention/efficient_attention/eva.py:138-233          (EVA, non-causal)
    efficient-attention/efficient_attention/lara.py:84-251          (LARA / LinearRA)
    efficient-attention/efficient_attention/causal_eva.py:458-788   (causal EVA, parallel branch)
    efficient-attention/efficient_attention/local_attention.py:134-182, abstract_attention.py:91-133
    efficient-attention/efficient_attention/attn_utils.py:12-30,155-234,292-348

It is *not* a transliteration: windows / chunks are expressed as explicit gather-index tables
(``-1`` = outside the sequence) instead of ``F.pad`` + ``as_strided`` + ``rearrange``, masks are
carried as index predicates, and the module-level wrappers take a flat ``state_dict``.

Pinning: the reference ships **no** golden vectors or tests for this path, so
the oracle is pinned against outputs of the *reference itself*, generated in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference`` with a 3-line ``timm`` shim) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every fixture.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch

MASK_VAL = -5.0e4  # eva.py:139, causal_eva.py:488, local_attention.py:142


# --------------------------------------------------------------------------------------------
# index tables: which token sits at slot l of group g (window or chunk); -1 = off the sequence
# --------------------------------------------------------------------------------------------
def group_index_1d(n: int, size: int, left: int, right: int) -> torch.Tensor:
    """[n // size, left + size + right] token ids (attn_utils.py:155-166, causal_eva.py:102-113)."""
    if left == 0 and right == 0:
        assert n % size == 0, "non-overlapping partition needs n % size == 0"
    groups = n // size
    idx = torch.arange(groups).unsqueeze(1) * size - left + torch.arange(left + size + right).unsqueeze(0)
    idx[(idx < 0) | (idx >= n)] = -1
    return idx


def group_index_2d(gh: int, gw: int, size: int, ext: int) -> torch.Tensor:
    """[(gh//size)*(gw//size), (size+2ext)^2]; groups row-major, slots row-major
    (attn_utils.py:172-210)."""
    if ext == 0:
        assert gh % size == 0 and gw % size == 0
    ny, nx, t = gh // size, gw // size, size + 2 * ext
    yy = (torch.arange(ny) * size - ext).view(ny, 1, 1, 1) + torch.arange(t).view(1, 1, t, 1)
    xx = (torch.arange(nx) * size - ext).view(1, nx, 1, 1) + torch.arange(t).view(1, 1, 1, t)
    ok = (yy >= 0) & (yy < gh) & (xx >= 0) & (xx < gw)
    idx = yy * gw + xx
    idx = torch.where(ok, idx, torch.full_like(idx, -1))
    return idx.reshape(ny * nx, t * t)


def take_groups(t: torch.Tensor, idx: torch.Tensor, fill: float = 0.0) -> torch.Tensor:
    """t [B,h,N,d] -> [B,h,G,L,d]; slots with idx < 0 are ``fill``."""
    B, h, _, d = t.shape
    G, L = idx.shape
    out = t[:, :, idx.clamp(min=0).reshape(-1)].reshape(B, h, G, L, d).clone()
    out[:, :, idx < 0] = fill
    return out


def take_mask_groups(mask: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """mask [B,N] bool -> [B,G,L] bool, True where padded or off-sequence (pad_val=1)."""
    B = mask.shape[0]
    G, L = idx.shape
    out = mask[:, idx.clamp(min=0).reshape(-1)].reshape(B, G, L).clone()
    out[:, idx < 0] = True
    return out


# --------------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------------
def linear(x, w, b=None):
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def layer_norm(x, g, b, eps: float = 1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def prm_logits(data: torch.Tensor, omega: torch.Tensor) -> torch.Tensor:
    """d^-1/2 <omega_c, x_j> - d^-1/2 |x_j|^2 / 2  ->  [..., c, j]   (attn_utils.py:324-336,347)."""
    s = data.shape[-1] ** -0.5
    return s * (omega @ data.transpose(-1, -2)) - s * 0.5 * (data * data).sum(-1).unsqueeze(-2)


def t5_bucket(rel: torch.Tensor, causal: bool, num_buckets: int, max_distance: int) -> torch.Tensor:
    """Bucket index of relative position rel = k_pos - q_pos (eva.py:31-55, causal_eva.py:62-86).
    The large-distance branch is evaluated in float32 like the reference so bucket edges agree."""
    n = -rel
    ret = torch.zeros_like(n)
    if not causal:
        num_buckets //= 2
        ret = ret + (n < 0).long() * num_buckets
        n = n.abs()
    else:
        n = n.clamp(min=0)
    max_exact = num_buckets // 2
    nf = n.to(torch.float32) / max_exact
    large = max_exact + (torch.log(nf) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).long()
    large = large.clamp(max=num_buckets - 1)
    return ret + torch.where(n < max_exact, n, large)


def t5_bias(table: torch.Tensor, L: int, J: int, causal: bool, num_buckets: int, max_distance: int,
            scale: float) -> torch.Tensor:
    """table [buckets, hb] -> dense bias [hb, L, J] = table[bucket(j - i)] * scale.  The halo offset
    between query and key coordinates is ignored, as in the reference."""
    rel = torch.arange(J).view(1, J) - torch.arange(L).view(L, 1)
    bucket = t5_bucket(rel, causal, num_buckets, max_distance)
    return table[bucket].permute(2, 0, 1) * scale


def t5_num_buckets(window: int, ext: int) -> int:
    return max(min(int((window + ext) / 2), 64), 16)  # eva.py:114, causal_eva.py:373


def adaptive_pool_bins(n_in: int, n_out: int):
    return [(int(math.floor(i * n_in / n_out)), int(math.ceil((i + 1) * n_in / n_out))) for i in range(n_out)]


# --------------------------------------------------------------------------------------------
# EVA core: (q, k, v) -> o     (non-causal: eva.py:151-227; causal: causal_eva.py:666-783)
# --------------------------------------------------------------------------------------------
def eva_core(q, k, v, *, seq_shape: Sequence[int], window: int, ext: int, chunk: int, chunk_ext: int,
             wq=None, bq=None, gq=None, betq=None, wk=None, bk=None, gk=None, betk=None,
             mu_coeff: float = 0.5, use_q: bool = True, pad_mask: Optional[torch.Tensor] = None,
             noise: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
             causal: bool = False, halo_right: Optional[bool] = None, mask_queries: bool = False,
             return_stats: bool = False):
    """q,k,v [B,h,N,d]; pad_mask [B,N] bool (True = pad) or None; noise [B,h,C,d] or None;
    bias [hb,L,J] (hb in {1,h}) already scaled, or None.  gq/gk None => no LayerNorm ('no-ln').
    use_q False => adaptive_proj == 'none' (mu = 0).  halo_right: windows extend to the right too
    (EVA) or only to the left (CausalEVAttention, whatever its `causal` flag); mask_queries: local
    logits of padded *queries* are masked as well (causal_eva.py:742-754)."""
    B, h, N, d = q.shape
    scale = d ** -0.5
    two_d = len(seq_shape) == 2
    if halo_right is None:
        halo_right = not causal
    if pad_mask is None:
        pad_mask = torch.zeros(B, N, dtype=torch.bool)
    pad_mask = pad_mask.to(torch.bool)
    if two_d:
        assert not causal
        gh, gw = seq_shape
        qi = group_index_2d(gh, gw, window, 0)
        ki = group_index_2d(gh, gw, window, ext)
        ci = group_index_2d(gh, gw, chunk, chunk_ext)
    else:
        qi = group_index_1d(N, window, 0, 0)
        ki = group_index_1d(N, window, ext, ext if halo_right else 0)
        ci = group_index_1d(N, chunk, chunk_ext, chunk_ext if halo_right else 0)
    W, L = qi.shape
    J = ki.shape[1]
    C = ci.shape[0]

    # ---- per-chunk statistics (eva.py:155-196) ----
    cm = take_mask_groups(pad_mask, ci)                        # [B,C,Jc]
    keep = (~cm).to(q.dtype).unsqueeze(1).unsqueeze(-1)        # [B,1,C,Jc,1]
    cq, ck, cv = (take_groups(t, ci) * keep for t in (q, k, v))
    k_bar = linear(ck.mean(-2), wk, bk)
    if gk is not None:
        k_bar = layer_norm(k_bar, gk, betk)
    if use_q:
        q_bar = linear(cq.mean(-2), wq, bq)
        if gq is not None:
            q_bar = layer_norm(q_bar, gq, betq)
        mu = mu_coeff * (q_bar + k_bar)
    else:
        mu = torch.zeros_like(k_bar)
    omega = mu if noise is None else mu + noise
    lg = prm_logits(ck, omega.unsqueeze(-2)).squeeze(-2)       # [B,h,C,Jc]
    lg = lg.masked_fill(cm.unsqueeze(1), MASK_VAL)
    beta = (torch.softmax(lg, -1).unsqueeze(-1) * cv).sum(-2)   # [B,h,C,d]

    # ---- joint local + chunk softmax (eva.py:200-227) ----
    wq_ = take_groups(q, qi)                                    # [B,h,W,L,d]
    wk_ = take_groups(k, ki)
    wv_ = take_groups(v, ki)
    r = scale * torch.einsum('bhwld,bhcd->bhwlc', wq_, k_bar)
    s = scale * torch.einsum('bhwld,bhwjd->bhwlj', wq_, wk_)
    if bias is not None:
        s = s + bias.unsqueeze(0).unsqueeze(2)
    km = take_mask_groups(pad_mask, ki).unsqueeze(1).unsqueeze(-2)     # [B,1,W,1,J]
    if mask_queries:
        qm = take_mask_groups(pad_mask, qi).unsqueeze(1).unsqueeze(-1)  # [B,1,W,L,1]
        s = s.masked_fill(qm | km, MASK_VAL)
    else:
        s = s.masked_fill(km, MASK_VAL)
    if causal:
        fut = torch.ones(L, J, dtype=torch.bool).triu(1 + ext)
        s = s.masked_fill(fut, MASK_VAL)
        # chunk c is visible to a query of chunk cq only if c < cq (causal_eva.py:725-739)
        q_chunk = (qi // chunk)                                          # [W,L]
        hide = torch.arange(C).view(1, 1, C) >= q_chunk.unsqueeze(-1)    # [W,L,C]
        r = r.masked_fill(hide, MASK_VAL)
    p = torch.softmax(torch.cat([s, r], -1), -1)
    o_w = torch.einsum('bhwlj,bhwjd->bhwld', p[..., :J], wv_) + torch.einsum('bhwlc,bhcd->bhwld', p[..., J:], beta)
    o = torch.zeros_like(q)
    o[:, :, qi.reshape(-1)] = o_w.reshape(B, h, W * L, d)
    if return_stats:
        return o, k_bar, beta
    return o


def local_core(q, k, v, *, seq_shape, window, ext, pad_mask=None, bias=None):
    """Pure local-window attention (local_attention.py:134-182)."""
    B, h, N, d = q.shape
    if pad_mask is None:
        pad_mask = torch.zeros(B, N, dtype=torch.bool)
    if len(seq_shape) == 2:
        qi = group_index_2d(seq_shape[0], seq_shape[1], window, 0)
        ki = group_index_2d(seq_shape[0], seq_shape[1], window, ext)
    else:
        qi = group_index_1d(N, window, 0, 0)
        ki = group_index_1d(N, window, ext, ext)
    wq_, wk_, wv_ = take_groups(q, qi), take_groups(k, ki), take_groups(v, ki)
    s = d ** -0.5 * torch.einsum('bhwld,bhwjd->bhwlj', wq_, wk_)
    if bias is not None:
        s = s + bias.unsqueeze(0).unsqueeze(2)
    s = s.masked_fill(take_mask_groups(pad_mask.bool(), ki).unsqueeze(1).unsqueeze(-2), MASK_VAL)
    o_w = torch.einsum('bhwlj,bhwjd->bhwld', torch.softmax(s, -1), wv_)
    o = torch.zeros_like(q)
    o[:, :, qi.reshape(-1)] = o_w.reshape(B, h, -1, d)
    return o


def softmax_core(q, k, v, pad_mask=None):
    """Dense softmax attention (abstract_attention.py:115-133); padded keys get -inf."""
    s = q.shape[-1] ** -0.5 * (q @ k.transpose(-1, -2))
    if pad_mask is not None:
        s = s.masked_fill(pad_mask.bool().unsqueeze(1).unsqueeze(2), float('-inf'))
    return torch.softmax(s, -1) @ v


# --------------------------------------------------------------------------------------------
# LARA core (lara.py:129-246)
# --------------------------------------------------------------------------------------------
def lara_landmarks_2d(q, k, v, gh, gw, n_lm, *, wq, bq, gq, betq, wk, bk, gk, betk, mixed: bool, vmixed: bool,
                      dense: bool = False):
    """AdaptiveAvgPool2d(sqrt n_lm) over the token grid, Linear+LN (absent for 'no-param-pool') per head ('light',
    lara.py:141-151) or over all channels at once ('dense', lara.py:131-139: channel = head * d + feature), optional
    landmark mixing (lara.py:157-174)."""
    B, h, N, d = q.shape
    side = int(math.sqrt(n_lm))
    by, bx = adaptive_pool_bins(gh, side), adaptive_pool_bins(gw, side)

    def pool(t):
        g = t.reshape(B, h, gh, gw, d)
        rows = [torch.stack([g[:, :, y0:y1, x0:x1].mean((2, 3)) for (x0, x1) in bx], 2) for (y0, y1) in by]
        return torch.stack(rows, 2).reshape(B, h, side * side, d)

    q_bar, k_bar = pool(q), pool(k)
    if wq is not None and dense:
        merge = lambda t: t.permute(0, 2, 1, 3).reshape(B, side * side, h * d)
        split = lambda t: t.reshape(B, side * side, h, d).permute(0, 2, 1, 3)
        q_bar = split(layer_norm(linear(merge(q_bar), wq, bq), gq, betq))
        k_bar = split(layer_norm(linear(merge(k_bar), wk, bk), gk, betk))
    elif wq is not None:
        q_bar = layer_norm(linear(q_bar, wq, bq), gq, betq)
        k_bar = layer_norm(linear(k_bar, wk, bk), gk, betk)
    if mixed:
        lg = d ** -0.5 * (k_bar @ k_bar.transpose(-1, -2))
        if vmixed:
            v_bar = pool(v)
            lg = lg + torch.log(torch.linalg.vector_norm(v_bar, ord=2, dim=-1) + 1e-4).unsqueeze(-2)
        k_bar = torch.softmax(lg, -1) @ k_bar
    return q_bar, k_bar


def lara_landmarks_1d(q, k, n_lm, *, wq=None, bq=None, gq=None, betq=None, wk=None, bk=None, gk=None, betk=None):
    """Segment means (lara.py:84-127). 'adaptive-1d' applies Linear+LN to every token first."""
    B, h, N, d = q.shape
    if wq is not None:
        q = layer_norm(linear(q, wq, bq), gq, betq)
        k = layer_norm(linear(k, wk, bk), gk, betk)
    if N <= n_lm:
        return q, k
    seg = N // n_lm
    if N % n_lm == 0:
        return q.reshape(B, h, n_lm, seg, d).mean(-2), k.reshape(B, h, n_lm, seg, d).mean(-2)
    n_short = (seg + 1) * n_lm - N

    def split(t):
        a = t[:, :, :n_short * seg].reshape(B, h, n_short, seg, d).mean(-2)
        b = t[:, :, n_short * seg:].reshape(B, h, n_lm - n_short, seg + 1, d).mean(-2)
        return torch.cat([a, b], -2)

    return split(q), split(k)


def lara_core(q, k, v, q_bar, k_bar, *, mis_type: str = 'mis-opt', alpha_coeff: float = 1.0,
              pad_mask=None, noise=None, sample_mode: str = 'single'):
    """SNIS estimator given landmarks (lara.py:182-246).  sample_mode: 'single' (omega = mu [+ noise]),
    'antithetic' (noise [B,h,C,d] -> [mu+e ; mu-e]), 'multi' (noise [B,h,2C,d])."""
    d = q.shape[-1]
    scale = d ** -0.5
    mu = q_bar + k_bar
    if noise is None:
        omega = mu
        rep = 1
    elif sample_mode == 'single':
        omega, rep = mu + noise, 1
    elif sample_mode == 'antithetic':
        omega, rep = torch.cat([mu + noise, mu - noise], -2), 2
    else:
        omega, rep = mu.repeat(1, 1, 2, 1) + noise, 2
    A = prm_logits(q, omega)                                    # [B,h,S,N]
    Bk = prm_logits(k, omega)
    if pad_mask is not None:
        Bk = Bk.masked_fill(pad_mask.bool().unsqueeze(1).unsqueeze(-2), float('-inf'))
    kv = torch.softmax(Bk, -1) @ v                              # [B,h,S,d]
    if mis_type == 'mis-opt':
        t = torch.softmax(scale * (q_bar @ q.transpose(-1, -2)), -1)
        mu_r = mu.repeat(1, 1, rep, 1)
        t = t.repeat(1, 1, rep, 1)
        Lm = prm_logits(mu_r, omega)                            # [B,h,S(omega),S(mu)]
        lp = torch.diagonal(Lm, dim1=-1, dim2=-2).unsqueeze(-1)
        bh = torch.exp(lp - torch.logsumexp(Lm, -1, keepdim=True))
        alpha = bh + alpha_coeff * (t - t.mean(-2, keepdim=True))
        log_alpha = torch.log(alpha.clamp(min=1e-8))
    elif mis_type == 'mis-bh':
        Lm = prm_logits(mu, omega)
        log_alpha = 0.0
        lp = torch.logsumexp(Lm, -1, keepdim=True)
    elif mis_type == 'mis-biased':
        Lm = prm_logits(mu, omega)
        log_alpha = (scale * (mu @ q.transpose(-1, -2))).repeat(1, 1, rep, 1)
        lp = torch.logsumexp(Lm, -1, keepdim=True)
    else:
        raise NotImplementedError(mis_type)
    logw = log_alpha + A + torch.logsumexp(Bk, -1, keepdim=True) - lp
    return torch.softmax(logw, -2).transpose(-1, -2) @ kv       # [B,h,N,d]


# --------------------------------------------------------------------------------------------
# module-level wrappers: state_dict + config + x -> y   (what the parity tests compare)
# --------------------------------------------------------------------------------------------
def _split_qkv(x, sd, heads):
    B = x.shape[0]
    C = x.shape[-1]
    y = linear(x.reshape(B, -1, C), sd['qkv.weight'], sd.get('qkv.bias'))
    y = y.reshape(B, -1, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    return y[0], y[1], y[2]


def _pad_tokens(x, mask, multiple):
    """Right-pad [B,N,C] to a multiple of the window and extend/create the mask (attn_utils.py:12-30)."""
    B, N, _ = x.shape
    rem = (-N) % multiple
    if mask is None:
        mask = torch.zeros(B, N, dtype=torch.bool)
    mask = mask.to(torch.bool)
    if rem:
        x = torch.cat([x, x.new_zeros(B, rem, x.shape[-1])], 1)
        mask = torch.cat([mask, torch.ones(B, rem, dtype=torch.bool)], 1)
    return x, mask


def _ada(sd, name, ln=True):
    if f'{name}.0.weight' not in sd:
        return None, None, None, None
    w, b = sd[f'{name}.0.weight'], sd[f'{name}.0.bias']
    if ln and f'{name}.1.weight' in sd:
        return w, b, sd[f'{name}.1.weight'], sd[f'{name}.1.bias']
    return w, b, None, None


def local_bias_from_state(sd, cfg, heads, L, J, scale):
    """Dense [hb,L,J] bias for the local logits (local_attention.py:70-79, eva.py:212-216)."""
    if cfg.get('use_t5_rpe'):
        nb = t5_num_buckets(cfg['window_size'], cfg['ext'])
        return t5_bias(sd['rel_pos_bias.relative_attention_bias.weight'], L, J, cfg.get('causal', False), nb,
                       cfg['window_size'] + cfg['ext'], scale)
    if cfg.get('use_rpe') and cfg['window_size'] > 0:
        tab = sd['local_relative_position_bias_table']
        if cfg['attn_2d']:
            return tab[sd['relative_position_index'].reshape(-1)].reshape(L, J, heads).permute(2, 0, 1)
        return tab
    return None


def eva_forward(sd: Dict[str, torch.Tensor], cfg: dict, x, pad_mask=None, noise=None):
    """EVA.forward (eva.py:138-233).  cfg keys: num_heads, window_size, attn_2d, overlap_window,
    adaptive_proj, num_landmarks, use_rpe, use_t5_rpe."""
    heads = cfg['num_heads']
    B, *shape, C = x.shape
    orig_n = int(math.prod(shape))
    w = cfg['window_size']
    ext = max(1, w // 2) if cfg.get('overlap_window') else 0
    cfg = dict(cfg, ext=ext)
    if cfg['attn_2d']:
        assert len(shape) == 2 and shape[0] % w == 0 and shape[1] % w == 0
        xf, mask = x.reshape(B, -1, C), (pad_mask if pad_mask is not None else None)
        N = orig_n
        chunk = int(math.sqrt(N // cfg['num_landmarks']))
        seq_shape = tuple(shape)
    else:
        xf, mask = _pad_tokens(x.reshape(B, -1, C), pad_mask, w)
        N = xf.shape[1]
        chunk = int(N // cfg['num_landmarks'])
        seq_shape = (N,)
    q, k, v = _split_qkv(xf, sd, heads)
    d = C // heads
    L = w * w if cfg['attn_2d'] else w
    J = (w + 2 * ext) ** 2 if cfg['attn_2d'] else w + 2 * ext
    bias = local_bias_from_state(sd, cfg, heads, L, J, d ** -0.5)
    ap = cfg.get('adaptive_proj', 'default')
    wq, bq, gq, betq = _ada(sd, 'adaptive_mu_q', ln=(ap != 'no-ln'))
    wk, bk, gk, betk = _ada(sd, 'adaptive_mu_k', ln=(ap != 'no-ln'))
    o = eva_core(q, k, v, seq_shape=seq_shape, window=w, ext=ext, chunk=chunk, chunk_ext=ext,
                 wq=wq, bq=bq, gq=gq, betq=betq, wk=wk, bk=bk, gk=gk, betk=betk,
                 mu_coeff=0.5, use_q=(ap != 'none'), pad_mask=mask, noise=noise, bias=bias)
    y = o.permute(0, 2, 1, 3).reshape((B,) + tuple(seq_shape) + (C,))
    y = linear(y, sd['proj.weight'], sd['proj.bias'])
    return y[..., :orig_n, :]            # eva.py:230-231 (slices W' in 2-D: a no-op, Appendix B-8)


def local_forward(sd, cfg, x, pad_mask=None):
    """LocalAttention.forward via MultiheadAttention.forward (abstract_attention.py:80-89)."""
    heads = cfg['num_heads']
    B, *shape, C = x.shape
    w = cfg['window_size']
    ext = max(1, w // 2) if cfg.get('overlap_window') else 0
    q, k, v = _split_qkv(x.reshape(B, -1, C), sd, heads)
    N = q.shape[2]
    d = C // heads
    if cfg['attn_2d']:
        side = int(math.sqrt(N))
        seq_shape, L, J = (side, side), w * w, (w + 2 * ext) ** 2
        mask = pad_mask
        qp, kp, vp = q, k, v
    else:
        rem = (-N) % w
        pad = lambda t: torch.cat([t, t.new_zeros(B, heads, rem, d)], 2) if rem else t
        qp, kp, vp = pad(q), pad(k), pad(v)
        mask = torch.zeros(B, N, dtype=torch.bool) if pad_mask is None else pad_mask.bool()
        if rem:
            mask = torch.cat([mask, torch.ones(B, rem, dtype=torch.bool)], 1)
        seq_shape, L, J = (N + rem,), w, w + 2 * ext
    bias = local_bias_from_state(sd, dict(cfg, ext=ext, use_t5_rpe=False), heads, L, J, d ** -0.5)
    o = local_core(qp, kp, vp, seq_shape=seq_shape, window=w, ext=ext, pad_mask=mask, bias=bias)[:, :, :N]
    y = o.transpose(1, 2).reshape((B,) + tuple(shape) + (C,))
    return linear(y, sd['proj.weight'], sd['proj.bias'])


def softmax_forward(sd, cfg, x, pad_mask=None):
    heads = cfg['num_heads']
    B, *shape, C = x.shape
    q, k, v = _split_qkv(x.reshape(B, -1, C), sd, heads)
    o = softmax_core(q, k, v, pad_mask)
    y = o.transpose(1, 2).reshape((B,) + tuple(shape) + (C,))
    return linear(y, sd['proj.weight'], sd['proj.bias'])


def lara_forward(sd, cfg, x, pad_mask=None, noise=None):
    """LinearRA.forward (lara.py:177-251).  cfg keys: pool_module_type ('light' default | 'dense'), num_heads,
    num_landmarks, proposal_gen, mis_type, alpha_coeff, use_antithetics, use_multisample."""
    heads = cfg['num_heads']
    B, *shape, C = x.shape
    q, k, v = _split_qkv(x.reshape(B, -1, C), sd, heads)
    gen = cfg.get('proposal_gen', 'pool')
    has_params = not gen.startswith('no-param-pool')
    pq = dict(zip(('wq', 'bq', 'gq', 'betq'), (sd.get('q_bar_gen.2.weight'), sd.get('q_bar_gen.2.bias'),
                                               sd.get('q_bar_gen.3.weight'), sd.get('q_bar_gen.3.bias'))))
    pk = dict(zip(('wk', 'bk', 'gk', 'betk'), (sd.get('k_bar_gen.2.weight'), sd.get('k_bar_gen.2.bias'),
                                               sd.get('k_bar_gen.3.weight'), sd.get('k_bar_gen.3.bias'))))
    if gen.startswith('adaptive-1d'):
        pq = dict(wq=sd['q_bar_gen.0.weight'], bq=sd['q_bar_gen.0.bias'], gq=sd['q_bar_gen.1.weight'], betq=sd['q_bar_gen.1.bias'])
        pk = dict(wk=sd['k_bar_gen.0.weight'], bk=sd['k_bar_gen.0.bias'], gk=sd['k_bar_gen.1.weight'], betk=sd['k_bar_gen.1.bias'])
    if not has_params:
        pq = dict(wq=None, bq=None, gq=None, betq=None)
        pk = dict(wk=None, bk=None, gk=None, betk=None)
    if len(shape) == 2:
        q_bar, k_bar = lara_landmarks_2d(q, k, v, shape[0], shape[1], cfg['num_landmarks'], **pq, **pk,
                                         mixed=gen.endswith('mixed'), vmixed=gen.endswith('-vmixed'),
                                         dense=cfg.get('pool_module_type', 'light') == 'dense')
    else:
        if pad_mask is not None:
            keep = (~pad_mask.bool()).to(q.dtype).unsqueeze(1).unsqueeze(-1)
            q, k, v = q * keep, k * keep, v * keep
        if gen.startswith('adaptive-1d'):
            q_bar, k_bar = lara_landmarks_1d(q, k, cfg['num_landmarks'], **pq, **pk)
        else:
            q_bar, k_bar = lara_landmarks_1d(q, k, cfg['num_landmarks'])
    mode = 'single'
    if noise is not None and cfg.get('use_multisample'):
        mode = 'multi'
    elif noise is not None and cfg.get('use_antithetics'):
        mode = 'antithetic'
    o = lara_core(q, k, v, q_bar, k_bar, mis_type=cfg.get('mis_type', 'mis-opt'),
                  alpha_coeff=cfg.get('alpha_coeff', 1.0), pad_mask=pad_mask, noise=noise, sample_mode=mode)
    y = o.transpose(1, 2).reshape((B,) + tuple(shape) + (C,))
    return linear(y, sd['proj.weight'], sd['proj.bias'])


def causal_eva_forward(sd, cfg, query, pad_mask=None, noise=None):
    """CausalEVAttention.forward, parallel branch, self-attention (causal_eva.py:488-536, 666-788).
    query [T,B,C] -> [T,B,C].  cfg keys: num_heads, window_size, overlap_window, chunk_size,
    num_chunks, causal, use_t5_rpe, adaptive_proj ('qk' | 'no-ln')."""
    heads = cfg['num_heads']
    x = query.transpose(0, 1)
    B, T, C = x.shape
    w = cfg['window_size']
    ext = max(1, w) if cfg.get('overlap_window') else 0
    xf, mask = _pad_tokens(x, pad_mask, w)
    N = xf.shape[1]
    d = C // heads

    def heads_of(name):
        return linear(xf, sd[f'{name}.weight'], sd.get(f'{name}.bias')).reshape(B, N, heads, d).transpose(1, 2)

    q, k, v = heads_of('q_proj'), heads_of('k_proj'), heads_of('v_proj')
    chunk = cfg['chunk_size'] if cfg.get('chunk_size') is not None else int(N // cfg['num_chunks'])
    assert chunk < N, "reference raises NameError when chunk_size >= N (SURVEY Appendix B-9)"
    bias = None
    if cfg.get('use_t5_rpe') and w > 0:
        bias = t5_bias(sd['rel_pos_bias.relative_attention_bias.weight'], w, w + ext, bool(cfg['causal']),
                       t5_num_buckets(w, ext), w + ext, d ** -0.5)
    ln = cfg.get('adaptive_proj', 'qk') == 'qk'
    wq, bq, gq, betq = _ada(sd, 'adaptive_mu_q', ln)
    wk, bk, gk, betk = _ada(sd, 'adaptive_mu_k', ln)
    o = eva_core(q, k, v, seq_shape=(N,), window=w, ext=ext, chunk=chunk, chunk_ext=0,
                 wq=wq, bq=bq, gq=gq, betq=betq, wk=wk, bk=bk, gk=gk, betk=betk, mu_coeff=1.0,
                 pad_mask=mask, noise=noise, bias=bias, causal=bool(cfg['causal']), halo_right=False,
                 mask_queries=True)
    y = o.permute(0, 2, 1, 3).reshape(B, N, C)
    y = linear(y, sd['out_proj.weight'], sd.get('out_proj.bias'))
    return y[:, :T].transpose(0, 1).contiguous()
