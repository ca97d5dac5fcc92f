"""`ScatterBrain` ('scatterbrain'): local window attention + random-feature attention over everything outside the window, in one joint
softmax (arXiv 2110.15343; reference scatterbrain_attention.py:46-180).

Same constructor surface as the reference (the union of `KernelizedAttention` and `LocalAttention`), same buffers / parameters.  The
core is `scatterbrain_forward` of libeva_sm100 (csrc/rfa_kernels.cu): per-feature maxima and global sums of phi(k) v over the
sequence, then one CTA per window.  Two configurations of the reference are not built, loudly:
  * proj_method != 'favorp': the reference itself fails there (`log_proj_k` is unbound, scatterbrain_attention.py:80-84);
  * overlap_window=True: the reference returns NaN (the zero-padded halo slots carry log-feature 0, their "local" log-sum-exp exceeds
    the global one and log_add_exp(.., mask=(1, -1)) takes the log of a negative number), so there is nothing to match.
"""
import math

import torch
import torch.nn.functional as F

from . import _abi
from .kernelized_attention import KernelizedAttention, recompute_fn
from .local_attention import LocalAttention


def scatterbrain_core_torch(q, k, v, proj, *, seq_shape, window, pad_mask, bias):
    """float32 restatement on [B, N, H, d] views (halo-free windows); what the backward differentiates."""
    from ._recompute import _groups_1d, _groups_2d, _take
    B, N, H, d = q.shape
    lowp = q.dtype if q.dtype != torch.float32 else None      # 16-bit activations: the big contractions run in that format
    m = proj.shape[1]
    dn = d ** -0.25

    def logf(x):
        dd = torch.einsum('bnhd,hjd->bnhj', dn * x, proj.to(x.dtype)).float()
        return dd - 0.5 * dn * dn * (x.float() * x.float()).sum(-1, keepdim=True) - 0.5 * math.log(m)
    lq, lk = logf(q), logf(k)
    mm = (lambda a, b: (a.to(lowp) @ b.to(lowp)).float()) if lowp is not None else (lambda a, b: a @ b)
    if pad_mask is not None:
        lk = lk.masked_fill(pad_mask.to(torch.bool).view(B, N, 1, 1), float('-inf'))
    idx = _groups_2d(seq_shape[0], seq_shape[1], window, 0, q.device) if len(seq_shape) == 2 else _groups_1d(N, window, 0, 0, q.device)
    G, L = idx.shape
    flat = idx.reshape(-1)
    take = lambda t: t.index_select(1, flat).view(B, G, L, H, t.shape[-1]).permute(0, 3, 1, 2, 4)      # [B, H, G, L, .]
    wq, wk, wv, wlq, wlk = take(q), take(k), take(v), take(lq), take(lk)
    mx = lk.amax(1).detach()                                                     # [B, H, m]
    pk = torch.exp(lk - mx.unsqueeze(1))                                         # [B, N, H, m]
    wpk = torch.exp(wlk - mx.view(B, H, 1, 1, m))
    num = mm(pk.permute(0, 2, 3, 1), v.permute(0, 2, 1, 3)).unsqueeze(2) - mm(wpk.transpose(-1, -2), wv)
    den = pk.sum(1).unsqueeze(2) - wpk.sum(-2)                                   # [B, H, G, m]
    kv_stats = num / den.unsqueeze(-1).clamp(min=1e-3)
    glse = torch.logsumexp(lk, 1).unsqueeze(2)
    llse = torch.logsumexp(wlk, -2)
    a = torch.maximum(glse, llse)
    nonlocal_lse = a + torch.log(torch.exp(glse - a) - torch.exp(llse - a) + 1e-5)
    s = d ** -0.5 * mm(wq, wk.transpose(-1, -2))
    if bias is not None:
        s = s + bias.float().view(1, H, 1, L, L)
    if pad_mask is not None:
        wmask = pad_mask.to(torch.bool).index_select(1, flat).view(B, 1, G, 1, L)
        s = s.masked_fill(wmask, float('-inf'))
    p = torch.softmax(torch.cat([s, wlq + nonlocal_lse.unsqueeze(-2)], -1), -1)
    o_w = mm(p[..., :L], wv) + mm(p[..., L:], kv_stats)                            # [B, H, G, L, d]
    o = torch.zeros(B, N, H, d, dtype=torch.float32, device=q.device)
    o = o.index_copy(1, flat, o_w.permute(0, 2, 3, 1, 4).reshape(B, G * L, H, d))
    return o.reshape(B, N, H * d)


class ScatterBrain(KernelizedAttention, LocalAttention):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.apply(self._init_weights)

    def forward(self, x, key_padding_mask=None):
        if self.proj_method != 'favorp':
            raise NotImplementedError("ScatterBrain: only proj_method='favorp' runs in the reference (scatterbrain_attention.py:80-84)")
        if self.ext_size > 0:
            raise NotImplementedError('ScatterBrain with overlap_window=True returns NaN in the reference; not built')
        if self.attn_drop.p > 0 and self.training:
            raise NotImplementedError('attention-probability dropout is not built into the sm_100a kernels')
        B, *seq_shape, C = x.shape
        orig_n = int(math.prod(seq_shape))
        w = self.window_size
        mask = key_padding_mask
        if self.attn_2d:
            assert len(seq_shape) == 2 and seq_shape[0] % w == 0 and seq_shape[1] % w == 0
            x_flat, shape = x.reshape(B, orig_n, C), tuple(seq_shape)
        else:
            x_flat = x.reshape(B, orig_n, C)
            rem = (-orig_n) % w
            if rem:   # the reference pads x itself (local_attention.py:114-132): padded tokens carry the qkv bias and are masked
                x_flat = F.pad(x_flat, (0, 0, 0, rem))
                full = torch.zeros(B, orig_n + rem, dtype=torch.bool, device=x.device)
                if mask is not None:
                    full[:, :orig_n] = mask.to(torch.bool)
                full[:, orig_n:] = True
                mask = full
            shape = (orig_n + rem,)
        q, k, v, packed = self._qkv_heads(x_flat)
        proj = self.get_proj_matrix(device=x.device, dtype=torch.float32)
        differentiable = torch.is_grad_enabled() and (packed.requires_grad or (self.use_rpe and self.local_relative_position_bias_table.requires_grad))
        bias = self._window_bias(differentiable=differentiable)
        tensors = (q, k, v, proj) + ((bias,) if bias is not None else ())
        out = recompute_fn(
            lambda q_, k_, v_, p_, b_=None: _abi.scatterbrain_forward(q_, k_, v_, seq_shape=shape, window=w, proj=p_, pad_mask=mask, bias=b_),
            lambda q_, k_, v_, p_, b_=None: scatterbrain_core_torch(q_, k_, v_, p_, seq_shape=shape, window=w, pad_mask=mask, bias=b_).to(v_.dtype),
            *tensors)
        y = self.proj(out.view((B,) + shape + (C,)))
        if not self.attn_2d:
            y = y[:, :orig_n]
        return self.proj_drop(y)

    @staticmethod
    def add_attn_specific_args(parent_parser, struct_name="attn_args", prefix=""):
        from . import add_nested_argument
        parent_parser = LocalAttention.add_attn_specific_args(parent_parser, struct_name=struct_name, prefix=prefix)
        parser = parent_parser.add_argument_group("Attention")
        p = prefix + "-" if len(prefix) > 1 else ""
        common = dict(struct_name=struct_name, prefix=prefix)
        add_nested_argument(parser, '--{}approx-attn-dim'.format(p), default=64, type=int, help='number of random features', **common)
        add_nested_argument(parser, '--{}proj-method'.format(p), default='favorp', type=str, help='which attention method is used for RFA', **common)
        add_nested_argument(parser, '--{}cos-weighting'.format(p), action='store_true', default=False, help='', **common)
        add_nested_argument(parser, '--{}sample-scheme'.format(p), default='default', type=str, **common)
        return parent_parser
