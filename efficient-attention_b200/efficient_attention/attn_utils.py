"""Host-side helpers of the drop-in package (the device work lives in csrc/)."""
import math

import torch
import torch.nn.functional as F
from torch import nn


def pad_to_multiple(tensor, multiple, dim=-2, value=0, create_mask=False):
    """Right-pad `dim` (negative index) up to a multiple; optionally return the [B, N] padding mask
    (reference attn_utils.py:12-30)."""
    assert dim < 0
    n = int(tensor.shape[dim])
    rem = (-n) % multiple
    if rem:
        tensor = F.pad(tensor, (0, 0) * (-1 - dim) + (0, rem), value=value)
    if not create_mask:
        return tensor
    mask = torch.zeros(tensor.shape[0], tensor.shape[-2], dtype=torch.bool, device=tensor.device)
    if rem:
        mask[:, -rem:] = True
    return tensor, mask


class FlattenTranspose(nn.Module):
    """[B, C, H, W] -> [B, H*W, C]; kept so LARA's `q_bar_gen.{2,3}` parameter names match."""

    def forward(self, x):
        return x.flatten(2).permute(0, 2, 1)


def t5_bucket_table(n_query, n_key, causal, num_buckets, max_distance):
    """LongTensor [n_query, n_key] of T5 relative-position buckets for rel = key - query
    (reference eva.py:31-55 / causal_eva.py:62-86).  float32 log like the reference."""
    rel = torch.arange(n_key).view(1, -1) - torch.arange(n_query).view(-1, 1)
    n = -rel
    base = torch.zeros_like(n)
    if causal:
        n = n.clamp(min=0)
    else:
        num_buckets //= 2
        base = (n < 0).long() * num_buckets
        n = n.abs()
    exact = num_buckets // 2
    coarse = exact + (torch.log(n.float() / exact) / math.log(max_distance / exact) * (num_buckets - exact)).long()
    coarse = coarse.clamp(max=num_buckets - 1)
    return base + torch.where(n < exact, n, coarse)
