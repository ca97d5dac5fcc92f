"""`MultiheadAttention`: qkv / proj Linear layers + dense softmax attention
(reference abstract_attention.py:41-140).  The Linear layers stay cuBLAS; QK^T-softmax-PV runs in
libeva_sm100 (`eva_window_attention` with a single window spanning the sequence)."""
import math

import torch
import torch.nn as nn

from . import _abi, _recompute


class MultiheadAttention(nn.Module):
    def __init__(self, dim, num_heads, fp32=False, qkv_bias=True, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv_bias = qkv_bias
        self.fp32 = fp32  # accepted and ignored, as in the reference

        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.zeros_(m.bias)
            nn.init.ones_(m.weight)
        elif isinstance(m, nn.Conv2d):
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
            m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
            if m.bias is not None:
                m.bias.data.zero_()

    # ---- helpers shared by the subclasses -------------------------------------------------------
    def _qkv_heads(self, x_flat):
        """x_flat [B, N, C] -> q, k, v as zero-copy [B, N, H, D] views of the packed qkv projection
        (the reference permutes to [B, H, N, D] views instead; abstract_attention.py:72-78)."""
        B, N, _ = x_flat.shape
        packed = self.qkv(x_flat).view(B, N, 3, self.num_heads, self.head_dim)
        return packed[:, :, 0], packed[:, :, 1], packed[:, :, 2], packed

    def proj_and_split_heads(self, x):
        """API-compatible with the reference: returns q, k, v as [B, H, N, D] views."""
        B, *seq_shape, C = x.shape
        q, k, v, _ = self._qkv_heads(x.reshape(B, -1, C))
        return q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)

    def _core(self, q, k, v, packed, key_padding_mask, seq_shape):
        if self.attn_drop.p > 0 and self.training:
            raise NotImplementedError('attention-probability dropout is not built into the sm_100a kernels')
        B, N, H, D = q.shape
        geometry = dict(seq_shape=(N,), window=N, ext=0, chunk=0, chunk_ext=0, mask_is_neg_inf=True)
        if _recompute.needs_grad(packed):
            return _recompute.window_core(q, k, v, geometry=geometry, pad_mask=key_padding_mask)
        return _abi.eva_window_attention(q, k, v, _abi.eva_geometry(q, **geometry), pad_mask=key_padding_mask)

    def forward(self, x, key_padding_mask=None):
        B, *seq_shape, C = x.shape
        q, k, v, packed = self._qkv_heads(x.reshape(B, -1, C))
        out = self._core(q, k, v, packed, key_padding_mask, seq_shape)   # [B, N, C]
        x = self.proj(out.view((B,) + tuple(seq_shape) + (C,)))
        return self.proj_drop(x)

    @staticmethod
    def add_attn_specific_args(parent_parser, struct_name="attn_args", prefix=""):
        from . import add_nested_argument
        parser = parent_parser.add_argument_group("Attention")
        flag_prefix = prefix + "-" if len(prefix) > 1 else ""
        add_nested_argument(parser, '--{}fp32'.format(flag_prefix), struct_name=struct_name, prefix=prefix,
                            default=False, action='store_true')
        return parent_parser
