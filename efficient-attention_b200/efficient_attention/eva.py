"""`EVA`: local-window attention + per-chunk control variates under one joint softmax
(reference eva.py:69-244).  Everything between the qkv projection and the output projection runs
in libeva_sm100 (`eva_forward`)."""
import math
import warnings

import numpy as np
import torch
from torch import nn

from . import _abi, _recompute
from .attn_utils import pad_to_multiple, t5_bucket_table
from .local_attention import LocalAttention


class T5RelativePositionBias(nn.Module):
    """Per-head bucketed relative-position bias (reference eva.py:15-65).  Holds the same
    `relative_attention_bias` Embedding; the bucket table is precomputed per (i, j) extent."""

    def __init__(self, scale, num_heads, causal=False, num_buckets=32, max_distance=128):
        super().__init__()
        self.scale = scale
        self.causal = causal
        self.num_buckets = num_buckets
        self.max_distance = max_distance
        self.relative_attention_bias = nn.Embedding(num_buckets, num_heads)
        self._buckets = {}

    def dense(self, n_query, n_key):
        """float [num_heads, n_query, n_key], already multiplied by `scale`."""
        key = (n_query, n_key, self.relative_attention_bias.weight.device)
        if key not in self._buckets:
            self._buckets[key] = t5_bucket_table(n_query, n_key, self.causal, self.num_buckets,
                                                 self.max_distance).to(key[2])
        return self.relative_attention_bias(self._buckets[key]).permute(2, 0, 1) * self.scale

    def forward(self, x):
        i, j = x.shape[-2:]
        return self.dense(i, j).unsqueeze(0).unsqueeze(2)


class EVA(LocalAttention):
    def __init__(self, adaptive_proj='default', num_landmarks=49, use_t5_rpe=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.adaptive_proj = adaptive_proj
        d = self.head_dim
        if adaptive_proj == 'default':
            self.adaptive_mu_q = nn.Sequential(nn.Linear(d, d), nn.LayerNorm(d))
            self.adaptive_mu_k = nn.Sequential(nn.Linear(d, d), nn.LayerNorm(d))
        elif adaptive_proj == 'no-ln':
            self.adaptive_mu_q = nn.Sequential(nn.Linear(d, d))
            self.adaptive_mu_k = nn.Sequential(nn.Linear(d, d))
        elif adaptive_proj == 'none':
            self.adaptive_mu_k = nn.Sequential(nn.Linear(d, d), nn.LayerNorm(d))
        else:
            raise ValueError("adaptive_proj must be 'default', 'no-ln' or 'none', got %r" % (adaptive_proj,))
        self.use_t5_rpe = use_t5_rpe
        self.num_landmarks = num_landmarks
        if self.use_rpe and self.use_t5_rpe:
            raise NotImplementedError("--use-rpe and --use-t5-rpe are mutually exclusive.")
        if self.use_rpe:
            warnings.warn("--use-rpe selects the window-table relative position bias; "
                          "--use-t5-rpe alone selects the T5-style bucketed bias instead.")
        if self.use_t5_rpe:
            self.rel_pos_bias = T5RelativePositionBias(
                self.scale, num_heads=self.num_heads, causal=False,
                num_buckets=max(min(int((self.window_size + self.ext_size) / 2), 64), 16),
                max_distance=self.window_size + self.ext_size)
        self.apply(self._init_weights)

    def _process_input(self, x, key_padding_mask):
        """2-D: validate only.  1-D: right-pad x to a multiple of the window and build / extend the
        padding mask (reference eva.py:119-136)."""
        B, *seq_shape, C = x.shape
        if self.attn_2d:
            assert len(seq_shape) == 2
            if self.window_size > 0:
                assert seq_shape[0] % self.window_size == 0 and seq_shape[1] % self.window_size == 0
        elif self.window_size > 0:
            if key_padding_mask is None:
                x, key_padding_mask = pad_to_multiple(x, self.window_size, dim=-2, create_mask=True)
            else:
                x = pad_to_multiple(x, self.window_size, dim=-2)
                key_padding_mask = pad_to_multiple(key_padding_mask, self.window_size, dim=-1, value=True)
            seq_shape = [x.shape[-2]]
        return x, key_padding_mask, seq_shape

    def _adaptive_params(self):
        """(w_q, b_q, ln_gain_q, ln_bias_q, w_k, b_k, ln_gain_k, ln_bias_k); absent parts are None."""
        def parts(seq):
            lin = seq[0]
            ln = seq[1] if len(seq) > 1 else None
            return (lin.weight, lin.bias, ln.weight if ln is not None else None, ln.bias if ln is not None else None)
        q = parts(self.adaptive_mu_q) if self.adaptive_proj != 'none' else (None, None, None, None)
        return q + parts(self.adaptive_mu_k)

    def _adaptive(self):
        params = self._adaptive_params()
        return _abi.memo(self, 'adaptive', params, lambda: _abi.adaptive(*params, mu_coeff=0.5))

    def _local_bias(self, differentiable=False):
        if self.use_t5_rpe:
            w, e = self.window_size, self.ext_size
            L, J = (w * w, (w + 2 * e) ** 2) if self.attn_2d else (w, w + 2 * e)
            if differentiable:
                return self.rel_pos_bias.dense(L, J).float().contiguous()
            table = self.rel_pos_bias.relative_attention_bias.weight
            return _abi.memo(self, 'bias', (table,), lambda: self.rel_pos_bias.dense(L, J).detach().float().contiguous())
        return self._window_bias(differentiable)

    def forward(self, x, key_padding_mask=None, noise=None):
        """x: [B, H', W', C] (attn_2d) or [B, N, C]; key_padding_mask [B, N], True = padding.
        `noise` (optional, [B, heads, chunks, head_dim]) overrides the N(0,1) draw of training mode
        so that two implementations can be compared on identical samples."""
        B, *seq_shape, C = x.shape
        orig_n = int(np.prod(seq_shape))
        x, key_padding_mask, seq_shape = self._process_input(x, key_padding_mask)
        N = int(np.prod(seq_shape))
        q, k, v, packed = self._qkv_heads(x.reshape(B, N, C))
        chunk = int(math.sqrt(N // self.num_landmarks)) if self.attn_2d else int(N // self.num_landmarks)
        if chunk <= 0:
            raise ValueError('num_landmarks=%d is larger than the sequence (%d tokens)' % (self.num_landmarks, N))
        geometry = dict(seq_shape=tuple(seq_shape), window=self.window_size, ext=self.ext_size, chunk=chunk, chunk_ext=self.ext_size)
        geom = _abi.eva_geometry(q, **geometry)
        if self.training and noise is None:
            noise = torch.randn(B, self.num_heads, _abi.num_chunks(geom), self.head_dim, dtype=torch.float32,
                                device=x.device)
        params = self._adaptive_params()
        if _recompute.needs_grad(packed, *params, *self._bias_sources()):
            # training (vit/engine.py:47-62): kernel forward, backward by recomputation (see _recompute.py)
            out = _recompute.eva_core(q, k, v, geometry=geometry, mu_coeff=0.5, params=params, pad_mask=key_padding_mask,
                                      noise=noise, bias=self._local_bias(differentiable=True), packed=packed)
        else:
            out = _abi.eva_forward(q, k, v, geom, self._adaptive(), pad_mask=key_padding_mask, noise=noise,
                                   bias=self._local_bias())
        x = self.proj(out.view((B,) + tuple(seq_shape) + (C,)))
        x = x[..., :orig_n, :]          # eva.py:230-231 (slices W' in 2-D: a no-op)
        return self.proj_drop(x)

    def _bias_sources(self):
        if self.use_t5_rpe:
            return (self.rel_pos_bias.relative_attention_bias.weight,)
        return (self.local_relative_position_bias_table,) if self.use_rpe else ()

    @staticmethod
    def add_attn_specific_args(parent_parser, struct_name="attn_args", prefix=""):
        from . import add_nested_argument
        parent_parser = LocalAttention.add_attn_specific_args(parent_parser, struct_name=struct_name, prefix=prefix)
        parser = parent_parser.add_argument_group("attention")
        p = prefix + "-" if len(prefix) > 1 else ""
        common = dict(struct_name=struct_name, prefix=prefix)
        add_nested_argument(parser, '--{}adaptive-proj'.format(p), default='default', type=str, **common)
        add_nested_argument(parser, '--{}num-landmarks'.format(p), default=49, type=int, **common)
        add_nested_argument(parser, '--{}use-t5-rpe'.format(p), action='store_true', default=False, **common)
        return parent_parser
