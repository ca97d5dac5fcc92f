"""Test helpers: build a drop-in module from a golden-fixture config."""
import warnings
from argparse import Namespace

import efficient_attention as ea


def build_module(cfg):
    kind = cfg['kind']
    base = dict(dim=cfg.get('dim'), num_heads=cfg['num_heads'], qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if kind == 'eva':
            return ea.AttentionFactory.build_attention('eva', dict(
                base, use_rpe=cfg['use_rpe'], window_size=cfg['window_size'], attn_2d=cfg['attn_2d'],
                overlap_window=cfg['overlap_window'], adaptive_proj=cfg['adaptive_proj'],
                num_landmarks=cfg['num_landmarks'], use_t5_rpe=cfg['use_t5_rpe']))
        if kind == 'local':
            return ea.AttentionFactory.build_attention('local', dict(
                base, use_rpe=cfg['use_rpe'], window_size=cfg['window_size'], attn_2d=cfg['attn_2d'],
                overlap_window=cfg['overlap_window']))
        if kind == 'softmax':
            return ea.AttentionFactory.build_attention('softmax', base)
        if kind == 'lara':
            return ea.AttentionFactory.build_attention('lara', dict(
                base, num_landmarks=cfg['num_landmarks'], kernel_size=None, proposal_gen=cfg['proposal_gen'],
                use_antithetics=cfg['use_antithetics'], use_multisample=cfg['use_multisample'],
                pool_module_type=cfg['pool_module_type'], mis_type=cfg['mis_type'], alpha_coeff=cfg['alpha_coeff']))
        if kind == 'causal_eva':
            return ea.CausalEVAttention(embed_dim=cfg['embed_dim'], num_heads=cfg['num_heads'], self_attention=True,
                                        attn_args=Namespace(
                                            adaptive_proj=cfg['adaptive_proj'], num_chunks=cfg['num_chunks'],
                                            chunk_size=cfg['chunk_size'], causal=cfg['causal'],
                                            use_t5_rpe=cfg['use_t5_rpe'], window_size=cfg['window_size'],
                                            overlap_window=cfg['overlap_window']))
        if kind == 'performer':
            return ea.AttentionFactory.build_attention('performer', dict(
                base, approx_attn_dim=cfg['approx_attn_dim'], proj_method=cfg['proj_method'], cos_weighting=cfg['cos_weighting'],
                sample_scheme=cfg['sample_scheme']))
        if kind == 'ra':
            return ea.AttentionFactory.build_attention('ra', dict(base, num_samples=cfg['num_samples']))
        if kind == 'scatterbrain':
            return ea.AttentionFactory.build_attention('scatterbrain', dict(
                base, use_rpe=cfg['use_rpe'], window_size=cfg['window_size'], attn_2d=cfg['attn_2d'],
                overlap_window=cfg['overlap_window'], approx_attn_dim=cfg['approx_attn_dim']))
    raise KeyError(kind)


def set_draws(module, cfg, a, device):
    """Hand the fixture's recorded random draws (training-mode projection, multinomial key index, noise) to the module's test hooks."""
    import torch
    if cfg['kind'] in ('performer', 'scatterbrain'):
        module._proj_override = a['proj'].to(device=device, dtype=torch.float32) if a.get('proj') is not None else None
    if cfg['kind'] == 'ra':
        k_ind = a['k_ind'].to(device) if a.get('k_ind') is not None else None
        noise = a['noise'].to(device=device, dtype=torch.float32) if a.get('noise') is not None else None
        module._draw_override = (k_ind, noise)


def run_module(module, cfg, a, device, dtype):
    """Forward a golden fixture's input through a drop-in module on `device`."""
    import torch
    x = a['x'].to(device=device, dtype=dtype)
    mask = a['mask'].to(device) if a['mask'] is not None else None
    noise = a['noise'].to(device=device, dtype=torch.float32) if a['noise'] is not None else None
    if cfg['kind'] in ('performer', 'ra', 'scatterbrain'):
        set_draws(module, cfg, a, device)
        with torch.no_grad():
            return module(x, mask)
    with torch.no_grad():
        if cfg['kind'] == 'causal_eva':
            return module(x, x, x, key_padding_mask=mask, noise=noise)[0]
        if cfg['kind'] in ('eva', 'lara'):
            return module(x, mask, noise=noise)
        return module(x, mask)
