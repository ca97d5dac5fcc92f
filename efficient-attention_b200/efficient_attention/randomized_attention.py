"""`RandomizedAttention` ('ra'): linear-complexity randomized attention of arXiv 2204.04667 (reference randomized_attention.py:12-63).

out_n = softmax_m(scale w_n . k_m - scale |k_m|^2 / 2) v_m with w_n = mu_n (+ Gaussian noise in training mode) and
mu_n = q_n + mean(k) (num_samples == 0) | q_n + E_pi[k] (num_samples == -1) | q_n + k[one index drawn from pi_n] (otherwise),
pi = softmax(scale q k^T).  The softmax-over-keys pass with the key-norm bias is `ra_forward` (csrc/rfa_kernels.cu); E_pi[k] is the
dense-softmax kernel of this library with v := k.  The draw from pi: for head_dim 64 / 16-bit activations `ra_sample` draws it by
Gumbel-max on tensor cores (argmax_m of scale q.k + Gumbel noise is distributed as pi: no [N, N] probabilities); otherwise the
probabilities are evaluated by library ops and `torch.multinomial` draws, as in the reference.  The padding mask is ignored, as in the
reference.
"""
import torch

from . import _abi
from .abstract_attention import MultiheadAttention
from .kernelized_attention import recompute_fn


def ra_core_torch(q, k, v, extra, noise, scale):
    """float32 restatement on [B, N, H, d] views (what the backward differentiates).  extra: [B, N, H, d] or [1, 1, H, d]."""
    lowp = q.dtype if q.dtype != torch.float32 else None      # 16-bit activations: the two [N, N] contractions run in that format
    q, k, v = (t.transpose(1, 2) for t in (q, k, v))
    w = q.float() + extra.float().transpose(1, 2)
    if noise is not None:
        w = w + noise
    mm = (lambda a, b: (a.to(lowp) @ b.to(lowp)).float()) if lowp is not None else (lambda a, b: a.float() @ b.float())
    kf = k.float()
    logits = scale * mm(w, k.transpose(-1, -2)) - 0.5 * scale * (kf * kf).sum(-1).unsqueeze(-2)
    o = mm(torch.softmax(logits, -1), v)
    return o.transpose(1, 2).reshape(o.shape[0], o.shape[2], -1)


class RandomizedAttention(MultiheadAttention):
    def __init__(self, num_samples=1, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.num_samples = num_samples
        self._draw_override = None          # test hook: (k_ind int64 [B, H, N] or None, noise float32 [B, H, N, d] or None)
        self.apply(self._init_weights)

    def _core(self, q, k, v, packed, key_padding_mask, seq_shape):
        B, N, H, D = q.shape
        k_ind, noise = self._draw_override if self._draw_override is not None else (None, None)
        differentiable = torch.is_grad_enabled() and packed.requires_grad
        if self.num_samples == 0:
            mode, extra = 'mean', None
        elif self.num_samples == -1:
            geometry = dict(seq_shape=(N,), window=N, ext=0, chunk=0, chunk_ext=0, mask_is_neg_inf=True)
            if differentiable:
                from . import _recompute
                extra = _recompute.window_core(q, k, k, geometry=geometry)
            else:
                extra = _abi.eva_window_attention(q, k, k, _abi.eva_geometry(q, **geometry))
            mode = 'given'
        else:
            mode, extra = 'gather', None
            if k_ind is None:
                with torch.no_grad():   # reference :38-40: one key per query whatever num_samples says
                    if _abi.ra_sample_supported(q):
                        # Gumbel-max on tensor cores: the same distribution without the [N, N] probabilities; the seed comes from
                        # PyTorch's CPU generator (torch.manual_seed governs it, no device synchronisation)
                        k_ind = _abi.ra_sample(q, k, seed=int(torch.randint(0, 2 ** 62, (1,)).item()))
                    else:
                        pi = torch.softmax(self.scale * torch.einsum('bnhd,bmhd->bhnm', q.float(), k.float()), dim=-1)
                        k_ind = torch.multinomial(pi.reshape(B * H * N, N), 1, replacement=True).reshape(B, H, N)
        if self.training and noise is None and self._draw_override is None:
            noise = torch.randn(B, H, N, D, device=q.device, dtype=torch.float32)
        scale = self.scale

        def extra_rows(k_, extra_):
            if mode == 'mean':
                return k_.float().mean(1, keepdim=True)
            if mode == 'given':
                return extra_.view(B, N, H, D)
            return torch.gather(k_, 1, k_ind.transpose(1, 2).unsqueeze(-1).expand(B, N, H, D))

        if mode == 'given':
            return recompute_fn(lambda q_, k_, v_, e_: _abi.ra_forward(q_, k_, v_, mode=mode, extra=e_, noise=noise),
                                lambda q_, k_, v_, e_: ra_core_torch(q_, k_, v_, extra_rows(k_, e_), noise, scale).to(v_.dtype), q, k, v, extra)
        return recompute_fn(lambda q_, k_, v_: _abi.ra_forward(q_, k_, v_, mode=mode, k_ind=k_ind, noise=noise),
                            lambda q_, k_, v_: ra_core_torch(q_, k_, v_, extra_rows(k_, None), noise, scale).to(v_.dtype), q, k, v)

    @staticmethod
    def add_attn_specific_args(parent_parser, struct_name="attn_args", prefix=""):
        from . import add_nested_argument
        parent_parser = MultiheadAttention.add_attn_specific_args(parent_parser, struct_name=struct_name, prefix=prefix)
        parser = parent_parser.add_argument_group("Attention")
        p = prefix + "-" if len(prefix) > 1 else ""
        add_nested_argument(parser, '--{}num-samples'.format(p), struct_name=struct_name, prefix=prefix, default=1, type=int,
                            help='number of random features')
        return parent_parser
