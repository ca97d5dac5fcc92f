"""CPU tests of the drop-in boundary: plugin surface, argparse plumbing, checkpoint compatibility,
host-side index logic, the C-ABI export list and loud failure without CUDA.  (`-m "not gpu"`)"""
import argparse
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, golden_names, load_golden
from helpers import build_module
from oracle import eva_oracle as O

import efficient_attention as ea
from efficient_attention import _abi
from efficient_attention.attn_utils import pad_to_multiple, t5_bucket_table


def test_registry_names_and_errors():
    # the reference registry, name for name (efficient_attention/__init__.py:53-62)
    assert set(ea.AttentionFactory.attn_dict) == {'performer', 'softmax', 'local', 'lara', 'ra', 'scatterbrain', 'eva', 'causal_eva'}
    with pytest.raises(KeyError):
        ea.AttentionFactory.build_attention('nope', {})
    with pytest.raises(TypeError):
        ea.AttentionFactory.build_attention('eva', dict(dim=64, num_heads=2, bogus=1))


def test_eva_cli_namespace_matches_reference_probe():
    # the namespace the DeiT EVA command line produces
    p = argparse.ArgumentParser()
    ea.AttentionFactory.add_attn_specific_args(p, 'eva')
    a = p.parse_args(['--use-rpe', '--window-size', '7', '--attn-2d'], namespace=ea.NestedNamespace())
    assert vars(a.attn_args) == dict(fp32=False, use_rpe=True, window_size=7, attn_2d=True, overlap_window=False,
                                     adaptive_proj='default', num_landmarks=49, use_t5_rpe=False)


def test_prefixed_arguments_land_in_named_struct():
    p = argparse.ArgumentParser()
    ea.AttentionFactory.add_attn_specific_args(p, 'eva', struct_name='attn_args_encoder', prefix='encoder-attn')
    ea.AttentionFactory.add_attn_specific_args(p, 'causal_eva', struct_name='attn_args_decoder', prefix='decoder-attn')
    a = p.parse_args(['--encoder-attn-window-size', '8', '--decoder-attn-chunk-size', '16', '--decoder-attn-causal'],
                     namespace=ea.NestedNamespace())
    assert a.attn_args_encoder.window_size == 8 and a.attn_args_encoder.num_landmarks == 49
    assert a.attn_args_decoder.chunk_size == 16 and a.attn_args_decoder.causal is True
    assert a.attn_args_decoder.window_size == 4


def test_lara_cli_defaults():
    p = argparse.ArgumentParser()
    ea.AttentionFactory.add_attn_specific_args(p, 'lara')
    a = p.parse_args([], namespace=ea.NestedNamespace())
    assert vars(a.attn_args) == dict(fp32=False, num_landmarks=49, kernel_size=None, pool_module_type='light',
                                     mis_type='mis-opt', proposal_gen='pool', use_antithetics=False,
                                     use_multisample=False, alpha_coeff=1.0)


def test_remove_argument():
    p = argparse.ArgumentParser()
    p.add_argument('--foo', default=1)
    p.add_argument('--bar', default=2)
    ea.remove_argument(p, '--foo')
    assert vars(p.parse_args([])) == {'bar': 2}


@pytest.mark.parametrize('name', golden_names())
def test_reference_checkpoints_load_strict(name):
    """state_dict keys / shapes are part of the boundary."""
    cfg, sd, _ = load_golden(name, dtype=torch.float32)
    m = build_module(cfg)
    m.load_state_dict(sd, strict=True)
    if 'relative_position_index' in sd:   # our index formula == the reference's buffer
        fresh = build_module(cfg)
        assert torch.equal(fresh.relative_position_index, sd['relative_position_index'])


def test_param_counts_match_survey():
    from argparse import Namespace
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        eva = ea.AttentionFactory.build_attention('eva', dict(dim=192, num_heads=3, use_rpe=True, window_size=7, attn_2d=True))
    lara = ea.AttentionFactory.build_attention('lara', dict(dim=384, num_heads=6, proposal_gen='pool-mixed'))
    causal = ea.CausalEVAttention(512, 8, self_attention=True, attn_args=Namespace(
        adaptive_proj='qk', num_chunks=None, chunk_size=256, causal=True, use_t5_rpe=True, window_size=256,
        overlap_window=False))
    count = lambda m: sum(p.numel() for p in m.parameters())
    assert (count(eva), count(lara), count(causal)) == (157091, 599936, 1059264)


def test_pad_to_multiple():
    x = torch.ones(2, 10, 4)
    y, m = pad_to_multiple(x, 8, dim=-2, create_mask=True)
    assert y.shape == (2, 16, 4) and m.shape == (2, 16) and m[:, 10:].all() and not m[:, :10].any()
    assert y[:, 10:].abs().sum() == 0
    y2, m2 = pad_to_multiple(torch.ones(2, 16, 4), 8, dim=-2, create_mask=True)
    assert y2.shape == (2, 16, 4) and not m2.any()
    km = pad_to_multiple(torch.zeros(2, 10, dtype=torch.bool), 8, dim=-1, value=True)
    assert km.shape == (2, 16) and km[:, 10:].all()


@pytest.mark.parametrize('causal', [False, True])
@pytest.mark.parametrize('geom', [(49, 49, 16, 7), (8, 12, 16, 12), (256, 256, 64, 256), (16, 32, 16, 32)])
def test_t5_bucket_table_matches_oracle(geom, causal):
    L, J, nb, md = geom
    rel = torch.arange(J).view(1, J) - torch.arange(L).view(L, 1)
    assert torch.equal(t5_bucket_table(L, J, causal, nb, md), O.t5_bucket(rel, causal, nb, md))


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'eva_sm100.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(?:int|const char\s*\*)\s+((?:eva|lara|rfa|ra|scatterbrain)_\w+)\s*\(', header))
    assert {'eva_forward', 'eva_chunk_stats', 'eva_window_attention', 'lara_forward', 'eva_last_error', 'rfa_forward', 'ra_forward',
            'scatterbrain_forward'} <= declared
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert _abi.load().eva_sm100_abi_version() == 4


def test_cabi_rejects_bad_geometry_without_touching_the_gpu():
    lib = _abi.load()
    g = _abi.EvaGeometry(2, 3, 196, 64, 2, 14, 14, 5, 0, 0, 2, 0, 0, 0, 0, _abi.EVA_F16)  # 14 % 5 != 0
    assert lib.eva_num_chunks(ctypes.byref(g)) == -22
    assert b'window' in lib.eva_last_error()
    g = _abi.EvaGeometry(2, 3, 196, 48, 2, 14, 14, 7, 0, 0, 2, 0, 0, 0, 0, _abi.EVA_F16)  # head_dim 48
    assert lib.eva_num_chunks(ctypes.byref(g)) == -95
    g = _abi.EvaGeometry(2, 3, 784, 64, 2, 28, 28, 7, 0, 0, 4, 0, 0, 0, 0, _abi.EVA_F16)
    assert lib.eva_num_chunks(ctypes.byref(g)) == 49


def test_no_cpu_fallback():
    cfg, sd, a = load_golden('eva_2d_noln', dtype=torch.float32)
    m = build_module(cfg)
    m.load_state_dict(sd)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m.eval()(a['x'])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'efficient-attention_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('SURVEY', ''), os.path.join(dirpath, f)


def test_memo_follows_parameter_updates():
    """_abi.memo caches tensors derived from parameters per module; the entry must be rebuilt after an in-place update
    (optimizer step, load_state_dict), after the parameter's storage was replaced (.data = ..., .to()), and must not be
    shared between modules."""
    lin = torch.nn.Linear(4, 4).eval()                          # the cache serves eval-mode modules only (see below)
    calls = []

    def build():
        calls.append(1)
        return lin.weight.detach().double().clone()

    a = _abi.memo(lin, 'w64', (lin.weight, lin.bias), build)
    b = _abi.memo(lin, 'w64', (lin.weight, lin.bias), build)
    assert a is b and len(calls) == 1
    with torch.no_grad():
        lin.weight.mul_(2.0)                                     # in place: version counter changes
    c = _abi.memo(lin, 'w64', (lin.weight, lin.bias), build)
    assert len(calls) == 2 and torch.equal(c, lin.weight.detach().double())
    lin.weight.data = torch.ones(4, 4)                           # storage replaced
    d = _abi.memo(lin, 'w64', (lin.weight, lin.bias), build)
    assert len(calls) == 3 and torch.equal(d, torch.ones(4, 4, dtype=torch.float64))
    lin.load_state_dict({'weight': torch.zeros(4, 4), 'bias': torch.zeros(4)})
    e = _abi.memo(lin, 'w64', (lin.weight, lin.bias), build)
    assert len(calls) == 4 and float(e.abs().sum()) == 0.0
    other = torch.nn.Linear(4, 4).eval()
    _abi.memo(other, 'w64', (other.weight,), lambda: calls.append(1) or other.weight.detach().clone())
    assert len(calls) == 5
    # writes through `.data` do not move `_version` (fairseq FP16 optimizer sync, EMA scripts): a training-mode module never
    # caches and drops its entry, so the first eval forward after training sees them; in pure eval mode the explicit hook does
    n = len(calls)
    lin.train()
    lin.weight.data.copy_(torch.full((4, 4), 3.0))
    f = _abi.memo(lin, 'w64', (lin.weight, lin.bias), build)
    assert len(calls) == n + 1 and float(f[0, 0]) == 3.0 and 'w64' not in _abi._MEMO.get(lin, {})
    lin.weight.data.copy_(torch.full((4, 4), 4.0))
    lin.eval()
    g = _abi.memo(lin, 'w64', (lin.weight, lin.bias), build)
    assert len(calls) == n + 2 and float(g[0, 0]) == 4.0
    lin.weight.data.copy_(torch.full((4, 4), 5.0))               # eval mode + .data write: invisible to the key ...
    assert float(_abi.memo(lin, 'w64', (lin.weight, lin.bias), build)[0, 0]) == 4.0
    import efficient_attention
    efficient_attention.invalidate_caches(lin)                   # ... until the caller says so
    assert float(_abi.memo(lin, 'w64', (lin.weight, lin.bias), build)[0, 0]) == 5.0
    # float32 contiguous parameters are handed to the kernels as views of their own storage: nothing to go stale
    assert _abi._f32(lin.weight).data_ptr() == lin.weight.data_ptr()
    # the cache holds ctypes structures with raw pointers: it must stay out of the module (deepcopy / pickle / state_dict)
    import copy
    import pickle
    assert set(lin.state_dict()) == {'weight', 'bias'} and not any('memo' in k for k in lin.__dict__)
    clone = pickle.loads(pickle.dumps(copy.deepcopy(lin)))
    assert torch.equal(clone.weight, lin.weight)


def test_bench_reference_arm_line_on_cpu():
    """`bench.py --impl reference` needs no GPU: it must print ONE JSON line with the contract's keys, `impl: reference`,
    a `cpu_baseline` describing the run and an `e2e` object that repeats the line's own value with zero copy bytes."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['unit'] == 'tokens/s' and d['value'] > 0 and 'workload' in d['config']
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] and d['e2e']['value'] == d['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0


def test_integration_md_ctypes_stub_matches_the_abi():
    """The ctypes stub printed in INTEGRATION.md must declare the same EvaGeometry / EvaAdaptive fields as the binding the
    package uses (a maintainer copying a stale stub makes the library read past the struct)."""
    import re
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = text[text.index('class EvaGeometry(ctypes.Structure)'):text.index('class EvaAdaptive(ctypes.Structure)')]
    names = re.findall(r'"([a-z_]+)"', block)
    assert names == [n for n, _ in _abi.EvaGeometry._fields_], names
    call = re.search(r'g = EvaGeometry\(([^)]*)\)', text).group(1)
    assert len(call.split(',')) == len(_abi.EvaGeometry._fields_)
    header = open(os.path.join(ROOT, 'include', 'eva_sm100.h')).read()
    for n, _ in _abi.EvaGeometry._fields_:
        assert re.search(r'\b%s\b' % n, header), n


def test_random_feature_modules_cli_and_errors():
    """'performer' / 'ra' / 'scatterbrain' argparse surface (kernelized_attention.py:322-330, randomized_attention.py:56-63,
    scatterbrain_attention.py:166-180) and the two configurations that do not run in the reference either."""
    p = argparse.ArgumentParser()
    p = ea.AttentionFactory.add_attn_specific_args(p, 'performer')
    ns = p.parse_args(['--approx-attn-dim', '32', '--proj-method', 'relu', '--cos-weighting'], namespace=ea.NestedNamespace())
    assert vars(ns.attn_args) == dict(fp32=False, approx_attn_dim=32, proj_method='relu', cos_weighting=True, sample_scheme='default')
    p = ea.AttentionFactory.add_attn_specific_args(argparse.ArgumentParser(), 'ra')
    assert vars(p.parse_args([], namespace=ea.NestedNamespace()).attn_args) == dict(fp32=False, num_samples=1)
    p = ea.AttentionFactory.add_attn_specific_args(argparse.ArgumentParser(), 'scatterbrain')
    got = vars(p.parse_args(['--window-size', '7', '--attn-2d'], namespace=ea.NestedNamespace()).attn_args)
    assert got == dict(fp32=False, use_rpe=False, window_size=7, attn_2d=True, overlap_window=False, approx_attn_dim=64,
                       proj_method='favorp', cos_weighting=False, sample_scheme='default')
    x = torch.zeros(1, 16, 64)
    sb = ea.AttentionFactory.build_attention('scatterbrain', dict(dim=64, num_heads=2, window_size=4, overlap_window=True))
    with pytest.raises(NotImplementedError, match='NaN'):
        sb(x)
    sb = ea.AttentionFactory.build_attention('scatterbrain', dict(dim=64, num_heads=2, window_size=4, proj_method='relu'))
    with pytest.raises(NotImplementedError, match='favorp'):
        sb(x)
    with pytest.raises(NotImplementedError):
        ea.AttentionFactory.build_attention('performer', dict(dim=64, num_heads=2, proj_method='bogus'))


def test_orthogonal_projection_init():
    """create_proj_matrix(ortho=True) (kernelized_attention.py:187-227): orthogonal directions inside each block of head_dim rows."""
    from efficient_attention.kernelized_attention import create_proj_matrix
    w = create_proj_matrix(3, 80, 32, ortho=True).double()
    assert w.shape == (3, 80, 32)
    unit = w / w.norm(dim=-1, keepdim=True)
    for blk in (unit[:, :32], unit[:, 32:64], unit[:, 64:]):
        gram = blk @ blk.transpose(-1, -2)
        assert torch.allclose(gram, torch.eye(blk.shape[1], dtype=torch.float64).expand_as(gram), atol=1e-5)


@pytest.mark.parametrize('name', golden_names(['perf_', 'ra_', 'sb_']))
def test_backward_restatements_match_reference_output(name):
    """The float32 PyTorch restatements the random-feature modules differentiate in training (`*_core_torch`) reproduce the
    reference's forward on the fixtures (CPU): what the backward differentiates is the reference's function."""
    from efficient_attention import kernelized_attention as KA, randomized_attention as RA, scatterbrain_attention as SB
    from oracle import eva_oracle as O
    cfg, sd, a = load_golden(name, dtype=torch.float32)
    H = cfg['num_heads']
    x = a['x']
    B, C = x.shape[0], x.shape[-1]
    shape = tuple(x.shape[1:-1])
    mask = a['mask']
    xf = x.reshape(B, -1, C)
    w = cfg.get('window_size')
    if cfg['kind'] == 'scatterbrain' and not cfg['attn_2d']:
        xf, mask = O._pad_tokens(xf, mask, w)
    N = xf.shape[1]
    packed = (xf @ sd['qkv.weight'].t() + sd['qkv.bias']).view(B, N, 3, H, C // H)
    q, k, v = packed[:, :, 0], packed[:, :, 1], packed[:, :, 2]
    proj = a.get('proj')
    if proj is None:
        proj = sd.get('eval_proj', sd.get('random_proj'))
    if cfg['kind'] == 'performer':
        method = cfg['proj_method']
        if method == 'mlp-fourier':
            pytest.skip('features by library ops')
        nu = (cfg['approx_attn_dim'] // (C // H)) // 2
        o = KA.linear_attention_torch(KA.features_torch(q, method, True, proj, nu), KA.features_torch(k, method, False, proj, nu), v,
                                      cfg['cos_weighting'], mask)
    elif cfg['kind'] == 'ra':
        ns = cfg['num_samples']
        if ns == 0:
            extra = k.mean(1, keepdim=True)
        elif ns == -1:
            pi = torch.softmax((C // H) ** -0.5 * torch.einsum('bnhd,bmhd->bhnm', q, k), -1)
            extra = torch.einsum('bhnm,bmhd->bnhd', pi, k)
        else:
            extra = torch.gather(k, 1, a['k_ind'].transpose(1, 2).unsqueeze(-1).expand(B, N, H, C // H))
        o = RA.ra_core_torch(q, k, v, extra, a['noise'], (C // H) ** -0.5)
    else:
        L = w * w if cfg['attn_2d'] else w
        bias = O.local_bias_from_state(sd, dict(cfg, ext=0, use_t5_rpe=False), H, L, L, 1.0)
        seq_shape = shape if cfg['attn_2d'] else (N,)
        o = SB.scatterbrain_core_torch(q, k, v, proj, seq_shape=seq_shape, window=w, pad_mask=mask, bias=bias)
    y = (o @ sd['proj.weight'].t() + sd['proj.bias']).view((B,) + ((N,) if len(shape) == 1 else shape) + (C,))
    if len(shape) == 1:
        y = y[:, :shape[0]]
    err = float((y.double() - a['y'].double()).norm() / a['y'].double().norm())
    assert err < 2e-5, (name, err)


def test_rfa_cabi_rejects_bad_geometry_without_touching_the_gpu():
    """The ABI v4 entry points validate before they enqueue anything (errno-style codes, message in eva_last_error)."""
    lib = _abi.load()
    sz = ctypes.c_size_t(0)
    g = _abi.RfaGeometry(2, 3, 196, 64, _abi.RFA_METHOD['favorp'], 64, 1, 0, 0, _abi.EVA_F16)
    assert lib.rfa_feature_dim(ctypes.byref(g)) == 64
    assert lib.rfa_forward_workspace_bytes(ctypes.byref(g), ctypes.byref(sz)) == 0 and sz.value % 256 == 0 and sz.value > 0
    g = _abi.RfaGeometry(2, 3, 196, 64, _abi.RFA_METHOD['fourier'], 24, 1, 0, 1, _abi.EVA_F32)      # cosFormer doubles 2 x 24
    assert lib.rfa_feature_dim(ctypes.byref(g)) == 96
    g = _abi.RfaGeometry(2, 3, 196, 60, _abi.RFA_METHOD['favorp'], 64, 1, 0, 0, _abi.EVA_F16)       # head_dim % 8
    assert lib.rfa_feature_dim(ctypes.byref(g)) == -95
    g = _abi.RfaGeometry(2, 3, 196, 64, 42, 64, 1, 0, 0, _abi.EVA_F16)                             # unknown feature map
    assert lib.rfa_feature_dim(ctypes.byref(g)) == -22 and b'method' in lib.eva_last_error()
    g = _abi.RfaGeometry(2, 3, 196, 128, _abi.RFA_METHOD['dpfp'], 0, 2, 0, 1, _abi.EVA_F16)         # 2 * 128 * 2 * 2 features
    assert lib.rfa_feature_dim(ctypes.byref(g)) == -95
    sg = _abi.SbGeometry(2, 3, 196, 64, 2, 14, 14, 5, 64, _abi.EVA_F16)                            # 14 % 5
    assert lib.scatterbrain_forward_workspace_bytes(ctypes.byref(sg), ctypes.byref(sz)) == -22
    sg = _abi.SbGeometry(2, 3, 1024, 64, 1, 1, 1024, 128, 64, _abi.EVA_F16)                        # windows of 128 tokens
    assert lib.scatterbrain_forward_workspace_bytes(ctypes.byref(sg), ctypes.byref(sz)) == -95
    sg = _abi.SbGeometry(2, 3, 196, 64, 2, 14, 14, 7, 64, _abi.EVA_F16)
    assert lib.scatterbrain_forward_workspace_bytes(ctypes.byref(sg), ctypes.byref(sz)) == 0 and sz.value > 0
    rg = _abi.RaGeometry(2, 3, 196, 64, 1, _abi.EVA_F16)
    assert lib.ra_forward_workspace_bytes(ctypes.byref(rg), ctypes.byref(sz)) == 0 and sz.value >= 2 * 3 * 64 * 4
    rg = _abi.RaGeometry(2, 3, 196, 64, 7, _abi.EVA_F16)                                           # unknown mode: no pointers are read
    assert lib.ra_forward(ctypes.byref(rg), None, None, None, None, None, None, None, None, 0, None) == -22


def test_graft_entry_build_runs_here():
    """The driver's "does it build" check: build() compiles (or finds up to date) the library, imports the package and loads it."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('graft_entry_under_test', os.path.join(ROOT, '__graft_entry__.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
