"""`LocalAttention`: block-local attention over 1-D / 2-D windows with optional halo and relative
position bias (reference local_attention.py:25-194).  EVA inherits its window geometry."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _abi, _recompute
from .abstract_attention import MultiheadAttention


class LocalAttention(MultiheadAttention):
    def __init__(self, use_rpe=False, window_size=2, attn_2d=False, overlap_window=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.window_size = window_size
        self.attn_2d = attn_2d
        self.use_rpe = use_rpe if window_size > 0 else False
        self.ext_size = max(1, window_size // 2) if overlap_window else 0
        if self.use_rpe:
            w, e = window_size, self.ext_size
            if attn_2d:
                # table size and index formula follow local_attention.py:43-62, including the row
                # multiplier 2e+w that makes distinct offsets collide.
                self.local_relative_position_bias_table = nn.Parameter(
                    torch.zeros(2 * (w + e - 1) * (2 * e + w + 1) + 1, self.num_heads))
                ky, kx = torch.meshgrid(torch.arange(-e, e + w), torch.arange(-e, e + w), indexing='ij')
                qy, qx = torch.meshgrid(torch.arange(w), torch.arange(w), indexing='ij')
                dy = qy.reshape(-1, 1) - ky.reshape(1, -1) + e + w - 1
                dx = qx.reshape(-1, 1) - kx.reshape(1, -1) + e + w - 1
                self.register_buffer("relative_position_index", dy * (2 * e + w) + dx)
            else:
                self.local_relative_position_bias_table = nn.Parameter(torch.zeros(self.num_heads, w, w + 2 * e))
            nn.init.trunc_normal_(self.local_relative_position_bias_table, std=.02)
        self.apply(self._init_weights)

    def _window_bias(self, differentiable=False):
        """Dense float32 [H, L, J] bias for the local logits (reference add_rel_pos_bias, :70-79).  `differentiable`: keep the
        autograd link to the table (training); otherwise a cached, detached copy."""
        if not self.use_rpe:
            return None
        table = self.local_relative_position_bias_table
        if not self.attn_2d:
            return table

        def gather(detach=True):
            L, J = self.relative_position_index.shape
            dense = table[self.relative_position_index.reshape(-1)].view(L, J, self.num_heads).permute(2, 0, 1)
            return dense.detach().float().contiguous() if detach else dense.float().contiguous()
        if differentiable:
            return gather(detach=False)
        return _abi.memo(self, 'window_bias', (table, self.relative_position_index), gather)

    def _core(self, q, k, v, packed, key_padding_mask, seq_shape):
        B, N, H, D = q.shape
        w, e = self.window_size, self.ext_size
        if self.attn_2d:
            side = int(math.sqrt(N))
            assert side * side == N
            shape, mask = (side, side), key_padding_mask
        else:
            rem = (-N) % w
            if rem:   # the reference pads q, k, v (not x) with zeros here (local_attention.py:150-156)
                packed_p = F.pad(packed, (0, 0, 0, 0, 0, 0, 0, rem))
                q, k, v = packed_p[:, :, 0], packed_p[:, :, 1], packed_p[:, :, 2]
                mask = torch.zeros(B, N + rem, dtype=torch.bool, device=q.device)
                if key_padding_mask is not None:
                    mask[:, :N] = key_padding_mask.to(torch.bool)
                mask[:, N:] = True
            else:
                mask = key_padding_mask
            shape = (N + rem,)
        geometry = dict(seq_shape=shape, window=w, ext=e, chunk=0, chunk_ext=0)
        if _recompute.needs_grad(packed, *((self.local_relative_position_bias_table,) if self.use_rpe else ())):
            out = _recompute.window_core(q, k, v, geometry=geometry, pad_mask=mask, bias=self._window_bias(differentiable=True))
        else:
            out = _abi.eva_window_attention(q, k, v, _abi.eva_geometry(q, **geometry), pad_mask=mask, bias=self._window_bias())
        return out[:, :N]

    @staticmethod
    def add_attn_specific_args(parent_parser, struct_name="attn_args", prefix=""):
        from . import add_nested_argument
        parent_parser = MultiheadAttention.add_attn_specific_args(parent_parser, struct_name=struct_name, prefix=prefix)
        parser = parent_parser.add_argument_group("Attention")
        p = prefix + "-" if len(prefix) > 1 else ""
        common = dict(struct_name=struct_name, prefix=prefix)
        add_nested_argument(parser, '--{}use-rpe'.format(p), action='store_true', default=False, **common)
        add_nested_argument(parser, '--{}window-size'.format(p), default=4, type=int, **common)
        add_nested_argument(parser, '--{}attn-2d'.format(p), action='store_true', default=False, **common)
        add_nested_argument(parser, '--{}overlap-window'.format(p), action='store_true', default=False, **common)
        return parent_parser
