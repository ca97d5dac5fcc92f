"""The drop-in package inside the reference's OWN model (`-m gpu`): `evit_tiny_p8` / `evit_tiny_p16` / `evit_small_p16` from
the reference's vit/models/efficient_vit.py (vendored unmodified under oracle/_ref by oracle/make_ref.py) built with this
package as `efficient_attention`, against the SAME model whose attention layers are replaced by the float64 CPU oracle
(call sites efficient_vit.py:112,118-121).

Tolerances (relative L2 of the logits against float64 on identical weights and inputs):
  float32 model                 <= 1e-4   (12 layers of fp32 kernels + cuBLAS fp32/tf32-off GEMMs)
  float16 model (.half())       <= 4e-3   ... and no worse than 1.5x the same model with the reference's dense softmax
                                          attention in fp16: a 12-layer fp16 network (48 GEMMs rounding to 11 bits) sits at
                                          1-3e-3 whatever its attention is; the attention CORE's own 1e-3 bound is checked
                                          layer by layer in test_gpu_parity.py
"""
import copy
import ctypes
import warnings

import pytest
import torch
from torch import nn

from conftest import rel_l2
from oracle import eva_oracle as O
from oracle import rfa_oracle as R
from oracle import ref_loader

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.have_ref(), reason='oracle/_ref missing: run python oracle/make_ref.py in the build container')]


def _dev():
    return torch.device('cuda', 0)


def _build(name, attn):
    vm = ref_loader.vit_models()
    import efficient_attention as ea
    assert 'efficient-attention_b200' in ea.__file__          # the reference model must be calling the PRODUCT package
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        torch.manual_seed(0)
        m = getattr(vm, name)(ref_loader.deit_args(attn, num_classes=100)).eval()
    # reference init is std .02 everywhere (near-uniform softmaxes); give the attention layers logits of order 1
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if '.attn.' in n_ and p.dim() == 2 and 'bias_table' not in n_:
                p.copy_(torch.randn(p.shape, generator=g) * (0.7 / p.shape[1] ** 0.5))
            elif 'bias_table' in n_:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
    return m


class _OracleAttn(nn.Module):
    """The float64 CPU oracle behind the attention interface of the reference model (`attn(x) -> y`)."""

    def __init__(self, mod, kind):
        super().__init__()
        self.sd = {k: (v.detach().double() if v.is_floating_point() else v.detach()) for k, v in mod.state_dict().items()}
        self.kind = kind
        if kind == 'eva':
            self.cfg = dict(num_heads=mod.num_heads, window_size=mod.window_size, attn_2d=True, overlap_window=False,
                            adaptive_proj=mod.adaptive_proj, num_landmarks=mod.num_landmarks, use_rpe=mod.use_rpe, use_t5_rpe=False)
        elif kind == 'performer':
            self.cfg = dict(num_heads=mod.num_heads, proj_method=mod.proj_method, approx_attn_dim=mod.approx_attn_dim,
                            cos_weighting=mod.cos_weighting)
        elif kind == 'scatterbrain':
            self.cfg = dict(num_heads=mod.num_heads, window_size=mod.window_size, attn_2d=mod.attn_2d, overlap_window=False,
                            use_rpe=mod.use_rpe)
        else:
            self.cfg = dict(num_heads=mod.num_heads, num_landmarks=mod.num_landmarks, proposal_gen=mod.proposal_gen,
                            mis_type=mod.mis_type, alpha_coeff=mod.alpha_coeff)

    def forward(self, x):
        if self.kind == 'eva':
            return O.eva_forward(self.sd, self.cfg, x)
        if self.kind == 'performer':
            return R.performer_forward(self.sd, self.cfg, x)
        if self.kind == 'scatterbrain':
            return R.scatterbrain_forward(self.sd, self.cfg, x)
        return O.lara_forward(self.sd, self.cfg, x)


def _oracle_model(model, kind):
    """Same reference model in float64 on the CPU; every block's attention is the oracle on that block's weights."""
    ref = copy.deepcopy(model).cpu().double()
    for blk in ref.blocks:
        blk.attn = _OracleAttn(blk.attn, kind)
    return ref


def _path_count(path):
    from efficient_attention import _abi
    lib = _abi.load()
    lib.eva_debug_path_count.restype = ctypes.c_int
    return lib.eva_debug_path_count(ctypes.c_int(path))


@pytest.mark.parametrize('name,attn,fast_paths', [('evit_tiny_p8', 'eva', (1, 3)), ('evit_tiny_p16', 'eva', (1, 3)),
                                                  ('evit_small_p16', 'lara', None), ('evit_tiny_p16', 'performer', None),
                                                  ('evit_tiny_p8', 'scatterbrain', None)])
def test_reference_vit_with_dropin_attention_matches_oracle(name, attn, fast_paths):
    model = _build(name, attn)
    torch.manual_seed(2)
    x = torch.randn(2, 3, 224, 224)
    torch.backends.cudnn.allow_tf32 = False            # the patch-embedding convolution would otherwise run in TF32 (1e-4 on its own)
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        want32 = _oracle_model(model, attn)(x.double())
        got32 = model.to(_dev())(x.to(_dev())).cpu()
        err32 = rel_l2(got32, want32)
        # fp16: both sides see the fp16-rounded weights and input; the oracle evaluates them in float64
        model16 = copy.deepcopy(model).half()
        want16 = _oracle_model(model16.float(), attn)(x.half().double())
        model16 = model16.half().to(_dev())
        before = [_path_count(p) for p in range(4)]
        got16 = model16(x.half().to(_dev())).float().cpu()
        after = [_path_count(p) for p in range(4)]
        err16 = rel_l2(got16, want16)
    assert err32 < 1e-4, (name, err32)
    if attn in ('performer', 'scatterbrain'):          # the tcgen05 random-feature kernels took the fp16 layers
        from efficient_attention import _abi
        assert _abi.rfa_tc_launches() >= len(model.blocks)
        if attn == 'scatterbrain':
            assert _abi.load().eva_debug_sb_tc_launches() >= len(model.blocks)
    if fast_paths is not None:     # every layer must have taken a tcgen05 path, none the CUDA-core kernels
        assert after[0] == before[0] and sum(after[p] - before[p] for p in fast_paths) == len(model.blocks), (before, after)
    assert not torch.isnan(got16).any()
    assert err16 < 4e-3, (name, err16)
    print(f'{name}+{attn}: logits rel-L2 fp32 {err32:.2e}, fp16 {err16:.2e}')


def test_fp16_model_error_is_the_networks_not_the_attentions():
    """Yardstick for the fp16 tolerance above: the same ViT with DENSE SOFTMAX attention (drop-in `softmax`, one window over
    the sequence) in fp16 against float64.  EVA's fp16 logit error must stay within 1.5x of it."""
    def run(attn):
        model = _build('evit_tiny_p16', attn)
        torch.manual_seed(2)
        x = torch.randn(2, 3, 224, 224)
        with torch.no_grad():
            m16 = copy.deepcopy(model).half()
            if attn == 'eva':
                want = _oracle_model(m16.float(), attn)(x.half().double())
            else:
                ref = copy.deepcopy(m16).float().double()
                for blk in ref.blocks:
                    sd = {k: v.detach().double() for k, v in blk.attn.state_dict().items()}
                    heads = blk.attn.num_heads
                    blk.attn = _Lambda(lambda t, sd=sd, heads=heads: O.softmax_forward(sd, dict(num_heads=heads), t))
                want = ref(x.half().double())
            got = m16.half().to(_dev())(x.half().to(_dev())).float().cpu()
        return rel_l2(got, want)
    e_eva, e_soft = run('eva'), run('softmax')
    print(f'fp16 logits rel-L2: eva {e_eva:.2e}, softmax {e_soft:.2e}')
    assert e_eva < max(1.5 * e_soft, 1e-3), (e_eva, e_soft)


class _Lambda(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


def test_deit_leg_of_the_bench_runs_and_graph_matches_eager():
    """bench.py's DeiT-tiny-p8 leg on a small batch: eager and CUDA-graph timings, graph replay bit-identical to eager."""
    import bench
    d = bench.deit_leg(_dev(), 1, batch=8, warm=2, steps=3)
    assert 'unavailable' not in d, d
    assert d['images_per_s'] > 0 and d['eager_ms'] > 0 and d['unit'] == 'images/s'
    assert 'graph_error' not in d, d.get('graph_error')
    assert d['graph_equals_eager'] is True and d['graph_ms'] > 0
