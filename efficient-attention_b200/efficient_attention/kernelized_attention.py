"""`KernelizedAttention` ('performer'): random-feature linear attention (reference kernelized_attention.py:185-330).

Same constructor, buffers / parameters (`eval_proj`, `random_proj`, `feature_proj.*`) and argparse flags as the reference.  The feature
maps and the two contractions phi(K)^T V, phi(Q) (phi(K)^T V) run in libeva_sm100 (`rfa_forward`, csrc/rfa_kernels.cu); only
'mlp-fourier' computes its features with library ops first (a learnable Linear over all features is a GEMM) and hands them over.
Training: the forward is the same kernel call; the backward differentiates a float32 PyTorch restatement (`performer_core_torch`).
"""
import math

import torch
from torch import nn

from . import _abi
from .abstract_attention import MultiheadAttention


def _orthogonal_gaussian(rows, cols, device=None, dtype=None):
    """Stacked orthogonal blocks with chi-distributed row norms (reference :205-227: QR of a Gaussian block, transposed; the last
    block truncated; rows rescaled by the norms of an independent Gaussian matrix)."""
    blocks, left = [], rows
    while left > 0:
        qmat, _ = torch.linalg.qr(torch.randn(cols, cols), mode='reduced')
        blocks.append(qmat.t()[:min(left, cols)])
        left -= cols
    norms = torch.randn(rows, cols).norm(dim=1)
    return (norms.unsqueeze(1) * torch.cat(blocks)).to(device=device, dtype=dtype)


def create_proj_matrix(num_heads, proj_dim, input_dim, ortho=False, seed=0, device=None, dtype=None):
    if ortho:
        return torch.stack([_orthogonal_gaussian(proj_dim, input_dim, device=device, dtype=dtype) for _ in range(num_heads)])
    return torch.randn(num_heads, proj_dim, input_dim, device=device, dtype=dtype)


class DeterministicLearnableFourierFeatures(nn.Module):
    """'mlp-fourier' (reference :155-178): cos / sin of a learnable projection, then Linear + ReLU."""

    def __init__(self, num_heads, dim, fourier_dim, std=0.02):
        super().__init__()
        self.dim = dim
        self.random_proj = nn.Parameter(std * torch.randn(num_heads, fourier_dim // 2, dim))
        self.phi = nn.Sequential(nn.Linear(fourier_dim, fourier_dim), nn.ReLU())

    def forward(self, x, random_proj=None, is_query=False):
        """x [B, N, H, d] -> [B, H, N, M]."""
        px = torch.einsum('bnhd,hjd->bhnj', x, self.random_proj.to(x.dtype))
        return self.phi(torch.cat([px.cos(), px.sin()], dim=-1) * (self.dim ** -0.5))


def features_torch(x, method, is_query, proj=None, nu=1):
    """Differentiable restatement of the feature maps (reference :12-113) on [B, N, H, d] -> [B, H, N, M], float32 except the
    projection einsum, which runs in the activations' 16-bit format when there is one (as the reference does under autocast)."""
    x = x.transpose(1, 2)
    d = x.shape[-1]
    dn = d ** -0.25
    if method in ('favorp', 'relu', 'fourier'):
        m = proj.shape[1]
        dd = torch.einsum('bhnd,hjd->bhnj', dn * x, proj.to(x.dtype)).float()
        x = x.float()
        half_sq = 0.5 * dn * dn * (x * x).sum(-1, keepdim=True)
        if method == 'favorp':
            stab = (dd.amax(-1, keepdim=True) if is_query else dd.amax((-1, -2), keepdim=True)).detach()
            return m ** -0.5 * torch.exp(dd - half_sq - stab) + 1e-4
        if method == 'relu':
            return torch.relu(m ** -0.5 * dd) + 1e-3
        h = torch.exp(half_sq - half_sq.amax(-2, keepdim=True).detach())
        return h * (m ** -0.5) * torch.cat([torch.sin(dd), torch.cos(dd)], -1)
    x = x.float()
    if method == 'dpfp':
        x2 = torch.cat([torch.relu(x), torch.relu(-x)], -1)
        return torch.cat([x2] * nu, -1) * torch.cat([x2.roll(shifts=j, dims=-1) for j in range(1, nu + 1)], -1)
    if method == 'relu-only':
        return torch.relu(x) + 0.1
    if method == 'sigmoid-only':
        return torch.sigmoid(x) + 0.1
    raise KeyError(method)


def linear_attention_torch(qf, kf, v, cos_weighting, pad_mask):
    """[B, H, N, M] features, v [B, N, H, d] -> [B, N, H*d] (reference :115-153, :311-319)."""
    v = v.float().transpose(1, 2)
    if pad_mask is not None:
        kf = kf.masked_fill(pad_mask.to(torch.bool).unsqueeze(1).unsqueeze(-1), 0.0)
    if cos_weighting:
        n = v.shape[-2]
        ang = (math.pi / 2) * torch.arange(n, dtype=torch.float32, device=v.device) / n
        c, s = torch.cos(ang).view(1, 1, n, 1), torch.sin(ang).view(1, 1, n, 1)
        qf, kf = torch.cat([qf * c, qf * s], -1), torch.cat([kf * c, kf * s], -1)
    kv = kf.transpose(-1, -2) @ v
    den = (qf * kf.sum(-2, keepdim=True)).sum(-1, keepdim=True)
    o = (qf @ kv) / den.clamp(min=1e-2)
    return o.transpose(1, 2).reshape(o.shape[0], o.shape[2], -1)


class _RfaFn(torch.autograd.Function):
    """Forward: `rfa_forward` (CUDA).  Backward: autograd through `torch_fn` re-evaluated on the saved inputs."""

    @staticmethod
    def forward(ctx, run_kernel, torch_fn, *tensors):
        ctx.torch_fn, ctx.n = torch_fn, len(tensors)
        ctx.save_for_backward(*tensors)
        with torch.no_grad():
            return run_kernel(*tensors)

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        with torch.enable_grad():
            ins = [t.detach().requires_grad_(t.is_floating_point() and need) for t, need in zip(saved, ctx.needs_input_grad[2:])]
            out = ctx.torch_fn(*ins)
            live = [t for t in ins if t.requires_grad]
            grads = torch.autograd.grad(out, live, grad_out.to(out.dtype))
        it = iter(grads)
        return (None, None) + tuple(next(it).to(t.dtype) if t.requires_grad else None for t in ins)


def recompute_fn(run_kernel, torch_fn, *tensors):
    """`run_kernel(*tensors)` now; if anything needs a gradient, `torch_fn(*tensors)` is what gets differentiated."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        return _RfaFn.apply(run_kernel, torch_fn, *tensors)
    return run_kernel(*tensors)


class KernelizedAttention(MultiheadAttention):
    def __init__(self, approx_attn_dim=64, proj_method='favorp', cos_weighting=False, sample_scheme='default', *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.approx_attn_dim = approx_attn_dim
        self.proj_method = proj_method
        self.cos_weighting = cos_weighting
        self.sample_scheme = sample_scheme
        self.use_random_proj = False
        self._proj_override = None          # test hook: the projection to use instead of a fresh training-mode draw
        if proj_method == 'dpfp':
            self.nu = (approx_attn_dim // self.head_dim) // 2
            assert self.nu > 0, "approx_attn_dim must be a multiple of 2*head_dim!"
        elif proj_method == 'mlp-fourier':
            self.feature_proj = DeterministicLearnableFourierFeatures(self.num_heads, self.head_dim, approx_attn_dim)
        elif proj_method in ('favorp', 'relu', 'fourier'):
            self.use_random_proj = True
            make = lambda: create_proj_matrix(self.num_heads, approx_attn_dim, self.head_dim, ortho=True)
            if sample_scheme == 'default':
                self.register_buffer('eval_proj', make())
            elif sample_scheme == 'fixed':
                self.register_buffer('random_proj', make())
            elif sample_scheme == 'learnable':
                self.random_proj = nn.Parameter(make())
            else:
                raise NotImplementedError('other sample schemes are not implemented yet.')
        elif proj_method not in ('relu-only', 'sigmoid-only'):
            raise NotImplementedError
        self.apply(self._init_weights)

    def get_proj_matrix(self, device=None, dtype=None):
        if not self.use_random_proj:
            return None
        if self.sample_scheme == 'default':
            if self._proj_override is not None:
                return self._proj_override.to(device=device)
            if self.training:
                return create_proj_matrix(self.num_heads, self.approx_attn_dim, self.head_dim, ortho=False, device=device, dtype=dtype)
            return self.eval_proj
        return self.random_proj

    def _core(self, q, k, v, packed, key_padding_mask, seq_shape):
        if self.attn_drop.p > 0 and self.training:
            raise NotImplementedError('attention-probability dropout is not built into the sm_100a kernels')
        method, cosw, mask = self.proj_method, self.cos_weighting, key_padding_mask
        if method == 'mlp-fourier':
            qf, kf = self.feature_proj(q.float(), is_query=True), self.feature_proj(k.float(), is_query=False)
            return recompute_fn(
                lambda qf_, kf_, v_: _abi.rfa_forward(q, k, v_, method='given', q_feat=qf_, k_feat=kf_, cos_weighting=cosw, pad_mask=mask),
                lambda qf_, kf_, v_: linear_attention_torch(qf_, kf_, v_, cosw, mask).to(v_.dtype), qf, kf, v)
        proj = self.get_proj_matrix(device=q.device, dtype=torch.float32)
        nu = getattr(self, 'nu', 1)
        if proj is None:
            return recompute_fn(
                lambda q_, k_, v_: _abi.rfa_forward(q_, k_, v_, method=method, nu=nu, cos_weighting=cosw, pad_mask=mask),
                lambda q_, k_, v_: linear_attention_torch(features_torch(q_, method, True, nu=nu), features_torch(k_, method, False, nu=nu),
                                                          v_, cosw, mask).to(v_.dtype), q, k, v)
        return recompute_fn(
            lambda q_, k_, v_, p_: _abi.rfa_forward(q_, k_, v_, method=method, proj=p_, cos_weighting=cosw, pad_mask=mask),
            lambda q_, k_, v_, p_: linear_attention_torch(features_torch(q_, method, True, p_), features_torch(k_, method, False, p_),
                                                          v_, cosw, mask).to(v_.dtype), q, k, v, proj)

    @staticmethod
    def add_attn_specific_args(parent_parser, struct_name="attn_args", prefix=""):
        from . import add_nested_argument
        parent_parser = MultiheadAttention.add_attn_specific_args(parent_parser, struct_name=struct_name, prefix=prefix)
        parser = parent_parser.add_argument_group("Attention")
        p = prefix + "-" if len(prefix) > 1 else ""
        common = dict(struct_name=struct_name, prefix=prefix)
        add_nested_argument(parser, '--{}approx-attn-dim'.format(p), default=64, type=int, help='number of random features', **common)
        add_nested_argument(parser, '--{}proj-method'.format(p), default='favorp', type=str, help='which random feature is used for RFA', **common)
        add_nested_argument(parser, '--{}cos-weighting'.format(p), action='store_true', default=False, help='', **common)
        add_nested_argument(parser, '--{}sample-scheme'.format(p), default='default', type=str, **common)
        return parent_parser
