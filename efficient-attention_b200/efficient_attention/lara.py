"""`LinearRA` (LARA): linear randomized attention with pooled landmark proposals and multiple
importance sampling (reference lara.py:14-268).  The core runs in libeva_sm100 (`lara_forward`)."""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _abi, _recompute
from .abstract_attention import MultiheadAttention
from .attn_utils import FlattenTranspose


class LinearRA(MultiheadAttention):
    def __init__(self, num_landmarks=49, kernel_size=None, proposal_gen='pool', use_antithetics=False,
                 use_multisample=False, pool_module_type='light', mis_type='mis-opt', alpha_coeff=1.0,
                 *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.num_landmarks = num_landmarks
        self.proposal_gen = proposal_gen
        self.use_antithetics = use_antithetics
        self.use_multisample = use_multisample
        self.pool_module_type = pool_module_type
        self.mis_type = mis_type
        self.alpha_coeff = alpha_coeff
        if pool_module_type == 'light':
            ch = self.head_dim
        elif pool_module_type == 'dense':
            ch = self.dim      # Linear / LayerNorm over all channels, i.e. across heads (lara.py:36-39)
        else:
            raise NotImplementedError(pool_module_type)
        if mis_type not in _abi.LARA_MIS:
            raise NotImplementedError("The attn_type {} is not supported yet.".format(mis_type))
        side = int(math.sqrt(num_landmarks))

        def pooled(with_params):
            mods = [nn.AdaptiveAvgPool2d(side), FlattenTranspose()]
            if with_params:
                mods += [nn.Linear(ch, ch), nn.LayerNorm(ch)]
            return nn.Sequential(*mods)

        if proposal_gen.startswith('pool'):
            self.q_bar_gen, self.k_bar_gen = pooled(True), pooled(True)
        elif proposal_gen.startswith('no-param-pool'):
            self.q_bar_gen, self.k_bar_gen = pooled(False), pooled(False)
        elif proposal_gen.startswith('adaptive-1d'):
            self.q_bar_gen = nn.Sequential(nn.Linear(ch, ch), nn.LayerNorm(ch))
            self.k_bar_gen = nn.Sequential(nn.Linear(ch, ch), nn.LayerNorm(ch))
        else:
            raise NotImplementedError(proposal_gen)
        self.apply(self._init_weights)

    def _proj_params(self, two_d):
        gen = self.proposal_gen
        if gen.startswith('adaptive-1d'):
            if two_d:
                raise ValueError("proposal_gen='adaptive-1d' expects [B, N, C] inputs")
            if self.pool_module_type == 'dense':     # the reference applies Linear(dim, dim) to [.., head_dim] rows here and fails
                raise ValueError("proposal_gen='adaptive-1d' needs pool_module_type='light'")
            lin_q, ln_q, lin_k, ln_k = self.q_bar_gen[0], self.q_bar_gen[1], self.k_bar_gen[0], self.k_bar_gen[1]
        elif gen.startswith('pool') and two_d and self.pool_module_type == 'light':
            lin_q, ln_q, lin_k, ln_k = self.q_bar_gen[2], self.q_bar_gen[3], self.k_bar_gen[2], self.k_bar_gen[3]
        else:  # 'no-param-pool', or pooled proposals on a 1-D input (plain segment means, lara.py:98-103)
            return _abi.adaptive(*([None] * 8), mu_coeff=1.0)
        params = (lin_q.weight, lin_q.bias, ln_q.weight, ln_q.bias, lin_k.weight, lin_k.bias, ln_k.weight, ln_k.bias)
        return _abi.memo(self, 'proj_params', params, lambda: _abi.adaptive(*params, mu_coeff=1.0, ln_eps=ln_q.eps))

    def _raw_proj_params(self, two_d):
        """The eight q_bar_gen / k_bar_gen parameters that take part in this call (None where absent) and the LayerNorm eps."""
        gen = self.proposal_gen
        if gen.startswith('adaptive-1d'):
            mods = (self.q_bar_gen[0], self.q_bar_gen[1], self.k_bar_gen[0], self.k_bar_gen[1])
        elif gen.startswith('pool') and two_d:
            mods = (self.q_bar_gen[2], self.q_bar_gen[3], self.k_bar_gen[2], self.k_bar_gen[3])
        else:
            return (None,) * 8, 1e-5
        lin_q, ln_q, lin_k, ln_k = mods
        return (lin_q.weight, lin_q.bias, ln_q.weight, ln_q.bias, lin_k.weight, lin_k.bias, ln_k.weight, ln_k.bias), ln_q.eps

    def _dense_landmarks(self, q, k, v, seq_shape, with_v):
        """pool_module_type == 'dense' (lara.py:131-139): AdaptiveAvgPool2d over the token grid of ALL channels, then Linear(dim, dim)
        + LayerNorm(dim) -- a library GEMM on [B, C, dim] -- and back to heads.  -> float32 [B, H, 3, C, d] for the kernels."""
        B, N, H, d = q.shape
        gh, gw = seq_shape
        side = int(math.sqrt(self.num_landmarks))

        def pooled(t):                         # [B, N, H, d] view -> [B, side^2, H * d]
            grid = t.reshape(B, gh, gw, H * d).permute(0, 3, 1, 2)
            return nn.functional.adaptive_avg_pool2d(grid.float(), side).flatten(2).transpose(1, 2)

        def proj(seq, m):
            lin, ln = seq[2], seq[3]
            y = nn.functional.linear(m, lin.weight.float(), None if lin.bias is None else lin.bias.float())
            return nn.functional.layer_norm(y, (y.shape[-1],), ln.weight.float(), ln.bias.float(), ln.eps)
        q_bar = proj(self.q_bar_gen, pooled(q)).view(B, -1, H, d)
        k_bar = proj(self.k_bar_gen, pooled(k)).view(B, -1, H, d)
        v_bar = pooled(v).view(B, -1, H, d) if with_v else torch.zeros_like(k_bar)
        return torch.stack([q_bar, k_bar, v_bar], 1).permute(0, 3, 1, 2, 4).contiguous()      # [B, H, 3, C, d]

    def forward(self, x, key_padding_mask=None, noise=None):
        """x: [B, H', W', C] or [B, N, C].  `noise` (test hook, not part of the reference signature) overrides the
        training-mode draw ([B, heads, C or 2C, head_dim], see lara.py:188-196).  As in the reference, the antithetic /
        multi-sample estimators exist in training mode only (lara.py:189-196 key on `self.training`): a `noise` passed to a
        module in eval mode is added as one sample per landmark and never switches the estimator."""
        B, *seq_shape, C = x.shape
        two_d = len(seq_shape) == 2
        if len(seq_shape) not in (1, 2):
            raise ValueError('expected [B, N, C] or [B, H, W, C]')
        N = int(np.prod(seq_shape))
        q, k, v, packed = self._qkv_heads(x.reshape(B, N, C))
        landmarks = int(math.sqrt(self.num_landmarks)) ** 2 if two_d else min(self.num_landmarks, N)
        mixed = 0
        if two_d and self.proposal_gen.endswith('-vmixed'):
            mixed = 2
        elif two_d and self.proposal_gen.endswith('mixed'):
            mixed = 1
        mode = _abi.LARA_SAMPLE_SINGLE
        if self.training or noise is not None:
            if self.training and self.use_multisample:
                mode, rows = _abi.LARA_SAMPLE_MULTI, 2 * landmarks
            elif self.training and self.use_antithetics:
                mode, rows = _abi.LARA_SAMPLE_ANTITHETIC, landmarks
            else:
                rows = landmarks
            if noise is None:
                noise = torch.randn(B, self.num_heads, rows, self.head_dim, dtype=torch.float32, device=x.device)
        given = None
        if two_d and self.pool_module_type == 'dense' and self.proposal_gen.startswith('pool'):
            given = self._dense_landmarks(q, k, v, seq_shape, mixed == 2)
        kernel_args = dict(seq_shape=tuple(seq_shape), landmarks=landmarks, per_token_proj=self.proposal_gen.startswith('adaptive-1d'),
                           mixed=mixed, mis_type=self.mis_type, sample_mode=mode,
                           zero_padded=(not two_d and key_padding_mask is not None), alpha_coeff=self.alpha_coeff)
        raw, ln_eps = self._raw_proj_params(two_d)
        if _recompute.needs_grad(packed, *raw):
            # training: kernel forward, backward by recomputation (see _recompute.py)
            out = _recompute.lara_core(q, k, v, kernel_args=kernel_args, params=raw, dense=given is not None, ln_eps=ln_eps,
                                       pad_mask=key_padding_mask, noise=noise, given=given)
        else:
            out = _abi.lara_forward(q, k, v, given_landmarks=given, proj=self._proj_params(two_d), pad_mask=key_padding_mask,
                                    noise=noise, **kernel_args)
        x = self.proj(out.view((B,) + tuple(seq_shape) + (C,)))
        return self.proj_drop(x)

    @staticmethod
    def add_attn_specific_args(parent_parser, struct_name="attn_args", prefix=""):
        from . import add_nested_argument
        parent_parser = MultiheadAttention.add_attn_specific_args(parent_parser, struct_name=struct_name, prefix=prefix)
        parser = parent_parser.add_argument_group("attention")
        p = prefix + "-" if len(prefix) > 1 else ""
        common = dict(struct_name=struct_name, prefix=prefix)
        add_nested_argument(parser, '--{}num-landmarks'.format(p), default=49, type=int, **common)
        add_nested_argument(parser, '--{}kernel-size'.format(p), default=None, type=int, **common)
        add_nested_argument(parser, '--{}pool-module-type'.format(p), default='light', type=str, **common)
        add_nested_argument(parser, '--{}mis-type'.format(p), default='mis-opt', type=str, **common)
        add_nested_argument(parser, '--{}proposal-gen'.format(p), default='pool', type=str, **common)
        add_nested_argument(parser, '--{}use-antithetics'.format(p), action='store_true', default=False, **common)
        add_nested_argument(parser, '--{}use-multisample'.format(p), action='store_true', default=False, **common)
        add_nested_argument(parser, '--{}alpha-coeff'.format(p), default=1.0, type=float, **common)
        return parent_parser
