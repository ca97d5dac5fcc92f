"""`CausalEVAttention`: causal EVA for fairseq decoders (reference causal_eva.py:296-914).

Parallel (training / scoring) branch only: chunk statistics, causal chunk mask, causal local window
and the joint softmax run in libeva_sm100.  Incremental decoding (causal_eva.py:542-665) is a later
round; passing `incremental_state` raises.  The fairseq protocol methods are kept so the module still
plugs into `TransformerDecoderLayerBase` (fairseq/modules/transformer_layer.py:295-321)."""
import math
import uuid
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import _abi, _recompute
from .attn_utils import pad_to_multiple, t5_bucket_table


class T5RelativePositionBias(nn.Module):
    """Single-table (head-shared) T5 bias (reference causal_eva.py:47-97)."""

    def __init__(self, scale, causal=False, num_buckets=32, max_distance=128):
        super().__init__()
        self.scale = scale
        self.causal = causal
        self.num_buckets = num_buckets
        self.max_distance = max_distance
        self.relative_attention_bias = nn.Embedding(num_buckets, 1)
        self._buckets = {}

    def dense(self, n_query, n_key):
        key = (n_query, n_key, self.relative_attention_bias.weight.device)
        if key not in self._buckets:
            self._buckets[key] = t5_bucket_table(n_query, n_key, self.causal, self.num_buckets,
                                                 self.max_distance).to(key[2])
        return self.relative_attention_bias(self._buckets[key]).squeeze(-1) * self.scale

    def forward(self, x):
        return self.dense(x.shape[-2], x.shape[-1])


class _Dropout(nn.Module):
    """Same interface as the FairseqDropout the reference vendors (causal_eva.py:218-251)."""

    def __init__(self, p, module_name=None):
        super().__init__()
        self.p = p
        self.module_name = module_name
        self.apply_during_inference = False

    def active(self):
        return self.p > 0 and (self.training or self.apply_during_inference)

    def make_generation_fast_(self, name: str, retain_dropout: bool = False, retain_dropout_modules=None, **kwargs):
        if retain_dropout and (retain_dropout_modules is None or self.module_name in retain_dropout_modules):
            self.apply_during_inference = True


class CausalEVAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, kdim=None, vdim=None, dropout=0.0, bias=True, self_attention=False,
                 q_noise=0.0, qn_block_size=8, attn_args=None):
        super().__init__()
        self._incremental_state_id = str(uuid.uuid4())
        self.embed_dim = embed_dim
        self.kdim = kdim if kdim is not None else embed_dim
        self.vdim = vdim if vdim is not None else embed_dim
        self.qkv_same_dim = self.kdim == embed_dim and self.vdim == embed_dim
        self.num_heads = num_heads
        self.dropout_module = _Dropout(dropout, module_name=self.__class__.__name__)
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        self.scaling = self.head_dim ** -0.5
        self.self_attention = self_attention
        assert not self.self_attention or self.qkv_same_dim, \
            "Self-attention requires query, key and value to be of the same size"
        if q_noise > 0:
            raise NotImplementedError('quant-noise (q_noise > 0) is outside the accelerated path')

        self.k_proj = nn.Linear(self.kdim, embed_dim, bias=bias)
        self.v_proj = nn.Linear(self.vdim, embed_dim, bias=bias)
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)

        self.window_size = attn_args.window_size
        self.ext_size = max(1, self.window_size) if attn_args.overlap_window else 0
        self.causal = attn_args.causal
        self.num_chunks = attn_args.num_chunks
        self.chunk_size = attn_args.chunk_size
        if self.chunk_size is not None:
            assert self.window_size >= self.chunk_size and self.window_size % self.chunk_size == 0
            self.num_chunks = None  # chunk_size overrides the number of landmarks
        self.use_t5_rpe = attn_args.use_t5_rpe if attn_args.window_size > 0 else False
        if self.use_t5_rpe:
            self.rel_pos_bias = T5RelativePositionBias(
                self.scaling, causal=self.causal,
                num_buckets=max(min(int((self.window_size + self.ext_size) / 2), 64), 16),
                max_distance=attn_args.window_size + self.ext_size)
        else:
            self.rel_pos_bias = None
        self.adaptive_proj = attn_args.adaptive_proj
        d = self.head_dim
        if self.adaptive_proj == 'qk':
            self.adaptive_mu_q = nn.Sequential(nn.Linear(d, d), nn.LayerNorm(d))
            self.adaptive_mu_k = nn.Sequential(nn.Linear(d, d), nn.LayerNorm(d))
        elif self.adaptive_proj == 'no-ln':
            self.adaptive_mu_q = nn.Sequential(nn.Linear(d, d))
            self.adaptive_mu_k = nn.Sequential(nn.Linear(d, d))
        else:
            raise NotImplementedError("adaptive_proj must be 'qk' or 'no-ln' (causal_eva.py:381-396,733)")
        self.reset_parameters()
        self.onnx_trace = False

    def prepare_for_onnx_export_(self):
        self.onnx_trace = True

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight, gain=1 / math.sqrt(2))
        elif isinstance(m, nn.LayerNorm):
            nn.init.zeros_(m.bias)
            nn.init.ones_(m.weight)

    def reset_parameters(self):
        gain = 1 / math.sqrt(2) if self.qkv_same_dim else 1.0
        for lin in (self.k_proj, self.v_proj, self.q_proj):
            nn.init.xavier_uniform_(lin.weight, gain=gain)
        self.adaptive_mu_q.apply(self._init_weights)
        self.adaptive_mu_k.apply(self._init_weights)
        nn.init.xavier_uniform_(self.out_proj.weight)
        if self.out_proj.bias is not None:
            nn.init.zeros_(self.out_proj.bias)

    def _process_input(self, x, key_padding_mask):
        if self.window_size > 0:
            if key_padding_mask is None:
                if x.shape[-2] % self.window_size == 0:
                    return x, None        # nothing padded: an all-False mask would only keep the kernels off their mask-free paths
                x, key_padding_mask = pad_to_multiple(x, self.window_size, dim=-2, create_mask=True)
            else:
                x = pad_to_multiple(x, self.window_size, dim=-2)
                key_padding_mask = pad_to_multiple(key_padding_mask, self.window_size, dim=-1, value=True)
        return x, key_padding_mask

    def _adaptive_params(self):
        def parts(seq):
            ln = seq[1] if len(seq) > 1 else None
            return (seq[0].weight, seq[0].bias, ln.weight if ln is not None else None, ln.bias if ln is not None else None)
        return parts(self.adaptive_mu_q) + parts(self.adaptive_mu_k)

    def _adaptive(self):
        params = self._adaptive_params()
        return _abi.memo(self, 'adaptive', params, lambda: _abi.adaptive(*params, mu_coeff=1.0))

    def _project_qkv_time_major(self, query):
        """Self-attention without padding and without autograd: q, k, v of all heads from ONE GEMM on the [T, B, C] input as it
        arrives (no [T,B,C] -> [B,T,C] copy, no three separate projections), returned as [B, T, H, D] strided views -- the
        kernels take any (batch, token, head) strides."""
        lins = (self.q_proj, self.k_proj, self.v_proj)
        srcs = tuple(l.weight for l in lins) + tuple(l.bias for l in lins)

        def fuse():
            w = torch.cat([l.weight.detach() for l in lins], 0).contiguous()
            b = None if lins[0].bias is None else torch.cat([l.bias.detach() for l in lins], 0).contiguous()
            return w, b
        w, b = _abi.memo(self, 'qkv_fused', srcs, fuse)
        T, B, C = query.shape
        qkv = F.linear(query, w, b).view(T, B, 3, self.num_heads, self.head_dim)
        return tuple(qkv[:, :, i].transpose(0, 1) for i in range(3))

    def forward(self, query, key: Optional[Tensor], value: Optional[Tensor],
                key_padding_mask: Optional[Tensor] = None,
                incremental_state: Optional[Dict[str, Dict[str, Optional[Tensor]]]] = None,
                need_weights: bool = True, attn_mask: Optional[Tensor] = None,
                noise: Optional[Tensor] = None, drop_mask: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
        """Time x Batch x Channel in and out; `attn_mask` is accepted and ignored like the reference
        (causality comes from the window / chunk masks).  Returns (output, None)."""
        if incremental_state is not None:
            raise NotImplementedError('incremental decoding is not built yet (SURVEY.md 8f-3)')
        time_major = query
        query = query.transpose(0, 1)
        bsz, tgt_len, embed_dim = query.size()
        assert embed_dim == self.embed_dim, f"query dim {embed_dim} != {self.embed_dim}"
        x, key_padding_mask = self._process_input(query, key_padding_mask)
        B, N, C = x.shape
        fused_qkv = ((self.self_attention or key is None) and N == tgt_len and not torch.is_grad_enabled()
                     and time_major.is_contiguous() and all((l.bias is None) == (self.q_proj.bias is None) for l in (self.k_proj, self.v_proj))
                     and self.kdim == self.embed_dim and self.vdim == self.embed_dim)
        if self.self_attention or key is None:
            k_in = v_in = x
        else:
            key, value = key.transpose(0, 1), value.transpose(0, 1)
            assert key.shape[0] == bsz and value is not None
            if key.shape[1] != tgt_len or value.shape[1] != tgt_len:
                raise NotImplementedError('CausalEVAttention needs key/value as long as the query')
            k_in = pad_to_multiple(key, self.window_size, dim=-2) if self.window_size > 0 else key
            v_in = pad_to_multiple(value, self.window_size, dim=-2) if self.window_size > 0 else value
        H, D = self.num_heads, self.head_dim
        if fused_qkv:
            q, k, v = self._project_qkv_time_major(time_major)
        else:
            q = self.q_proj(x).view(B, N, H, D)
            k = self.k_proj(k_in).view(B, N, H, D)
            v = self.v_proj(v_in).view(B, N, H, D)
        chunk = self.chunk_size if self.chunk_size is not None else int(N // self.num_chunks)
        if chunk >= N:
            raise ValueError('chunk size %d must be smaller than the padded sequence %d (causal_eva.py:680-683)' % (chunk, N))
        geometry = dict(seq_shape=(N,), window=self.window_size, ext=self.ext_size, chunk=chunk, chunk_ext=0,
                        causal=bool(self.causal), halo_left_only=True, mask_queries=True, bias_toeplitz=bool(self.use_t5_rpe))
        geom = _abi.eva_geometry(q, **geometry)
        if self.training and noise is None:
            noise = torch.randn(B, H, _abi.num_chunks(geom), D, dtype=torch.float32, device=x.device)
        table = self.rel_pos_bias.relative_attention_bias.weight if self.use_t5_rpe else None
        params = self._adaptive_params()
        dropping = self.dropout_module.active()
        grad = _recompute.needs_grad(q, k, v, table, *params)
        bias = None
        if self.use_t5_rpe:
            if grad or dropping:
                bias = self.rel_pos_bias.dense(self.window_size, self.window_size + self.ext_size).unsqueeze(0).float().contiguous()
            else:
                bias = _abi.memo(self, 'bias', (table,), lambda: self.rel_pos_bias.dense(
                    self.window_size, self.window_size + self.ext_size).unsqueeze(0).detach().float().contiguous())
        if dropping:
            # attention-probability dropout (causal_eva.py:778): the kernels have none, so this one configuration evaluates the core
            # with the float32 PyTorch recomputation in the forward as well (library ops on the device; autograd does the rest)
            wq_, bq_, gq_, betq_, wk_, bk_, gk_, betk_ = params
            out = _recompute.eva_core_torch(q, k, v, seq_shape=(N,), window=self.window_size, ext=self.ext_size, chunk=chunk,
                                            chunk_ext=0, wq=wq_, bq=bq_, gq=gq_, betq=betq_, wk=wk_, bk=bk_, gk=gk_, betk=betk_,
                                            mu_coeff=1.0, pad_mask=key_padding_mask, noise=noise, bias=bias, causal=bool(self.causal),
                                            left_only=True, mask_queries=True, p_drop=float(self.dropout_module.p),
                                            drop_mask=drop_mask).to(q.dtype)
        elif grad:
            out = _recompute.eva_core(q, k, v, geometry=geometry, mu_coeff=1.0, params=params, pad_mask=key_padding_mask,
                                      noise=noise, bias=bias)
        else:
            out = _abi.eva_forward(q, k, v, geom, self._adaptive(), pad_mask=key_padding_mask, noise=noise, bias=bias)
        y = self.out_proj(out)
        if tgt_len != N:
            y = y[:, :tgt_len]
        return y.transpose(0, 1).contiguous(), None

    # ---- fairseq incremental-state protocol (state layout of causal_eva.py:253-294, 835-866) -----
    def init_incremental_state(self):
        self._incremental_state_id = str(uuid.uuid4())

    def _get_full_incremental_state_key(self, key: str) -> str:
        return "{}.{}".format(self._incremental_state_id, key)

    def get_incremental_state(self, incremental_state, key):
        full_key = self._get_full_incremental_state_key(key)
        if incremental_state is None or full_key not in incremental_state:
            return None
        return incremental_state[full_key]

    def set_incremental_state(self, incremental_state, key, value):
        if incremental_state is not None:
            incremental_state[self._get_full_incremental_state_key(key)] = value
        return incremental_state

    def reorder_incremental_state(self, incremental_state, new_order: Tensor):
        buf = self._get_input_buffer(incremental_state)
        if buf is not None:
            for name, t in buf.items():
                if t is not None:
                    buf[name] = t.index_select(0, new_order)
            incremental_state = self._set_input_buffer(incremental_state, buf)
        return incremental_state

    def _get_input_buffer(self, incremental_state):
        result = self.get_incremental_state(incremental_state, "attn_state")
        return result if result is not None else {}

    def _set_input_buffer(self, incremental_state, buffer):
        return self.set_incremental_state(incremental_state, "attn_state", buffer)

    def apply_sparse_mask(self, attn_weights, tgt_len: int, src_len: int, bsz: int):
        return attn_weights

    def upgrade_state_dict_named(self, state_dict, name):
        """Split a legacy fused `in_proj_{weight,bias}` into q/k/v projections (causal_eva.py:871-900)."""
        prefix = name + "." if name != "" else ""
        for key in [k for k in state_dict.keys() if k.endswith(prefix + "in_proj_weight")]:
            w = state_dict.pop(key)
            dim = w.shape[0] // 3
            for i, part in enumerate(('q_proj', 'k_proj', 'v_proj')):
                state_dict[prefix + part + ".weight"] = w[i * dim:(i + 1) * dim]
            bias_key = prefix + "in_proj_bias"
            if bias_key in state_dict:
                b = state_dict.pop(bias_key)
                for i, part in enumerate(('q_proj', 'k_proj', 'v_proj')):
                    state_dict[prefix + part + ".bias"] = b[i * dim:(i + 1) * dim]

    @staticmethod
    def add_attn_specific_args(parent_parser, struct_name="attn_args", prefix=""):
        from . import add_nested_argument
        parser = parent_parser.add_argument_group("attention")
        p = prefix + "-" if len(prefix) > 1 else ""
        common = dict(struct_name=struct_name, prefix=prefix)
        add_nested_argument(parser, '--{}adaptive-proj'.format(p), default='default', type=str, **common)
        add_nested_argument(parser, '--{}num-chunks'.format(p), default=None, type=int, **common)
        add_nested_argument(parser, '--{}chunk-size'.format(p), default=None, type=int, **common)
        add_nested_argument(parser, '--{}causal'.format(p), action='store_true', default=False, **common)
        add_nested_argument(parser, '--{}use-t5-rpe'.format(p), action='store_true', default=False, **common)
        add_nested_argument(parser, '--{}window-size'.format(p), default=4, type=int, **common)
        add_nested_argument(parser, '--{}overlap-window'.format(p), action='store_true', default=False, **common)
        return parent_parser
