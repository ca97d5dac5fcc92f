"""Import helpers for `oracle/_ref/` (the vendored, unmodified reference; see oracle/make_ref.py).

TEST INFRASTRUCTURE, NOT PRODUCT.  Two ways in:

* `reference_attention()` -- puts oracle/_ref FIRST on sys.path and imports the reference's own `efficient_attention`.  Only valid
  in a process that has not imported the drop-in package of the same name (bench.py --impl reference, tests run in a subprocess).
* `vit_models()` -- imports the reference's `vit/models` package (as `ref_vit_models`) with oracle/_ref LAST on sys.path, so that
  its `from efficient_attention import AttentionFactory` binds to whatever `efficient_attention` the process already uses: the
  drop-in package in tests / bench (the reference model is the CALLER of the product), or the reference itself after
  `reference_attention()`.
"""
import importlib
import importlib.util
import os
import sys
from argparse import Namespace

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')


def have_ref():
    return os.path.isfile(os.path.join(REF, 'efficient_attention', '__init__.py'))


def reference_attention():
    if not have_ref():
        raise RuntimeError('oracle/_ref is missing: run `python oracle/make_ref.py` in the build container')
    mod = sys.modules.get('efficient_attention')
    if mod is not None and not os.path.abspath(mod.__file__).startswith(REF):
        raise RuntimeError('another `efficient_attention` is already imported in this process; the reference needs its own process')
    if REF not in sys.path:
        sys.path.insert(0, REF)
    mod = importlib.import_module('efficient_attention')
    assert os.path.abspath(mod.__file__).startswith(REF), mod.__file__
    return mod


def vit_models():
    if not have_ref():
        raise RuntimeError('oracle/_ref is missing: run `python oracle/make_ref.py` in the build container')
    if 'ref_vit_models' in sys.modules:
        return sys.modules['ref_vit_models']
    if REF not in sys.path:
        sys.path.append(REF)                       # last: only `timm` (shim) is resolved through it
    pkg_dir = os.path.join(REF, 'models')
    spec = importlib.util.spec_from_file_location('ref_vit_models', os.path.join(pkg_dir, '__init__.py'),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules['ref_vit_models'] = mod
    spec.loader.exec_module(mod)
    return mod


def deit_args(attn_name='eva', num_classes=1000, input_size=224, **attn_kw):
    """The argparse namespace vit/main.py would hand to `evit_*` for the README's DeiT + EVA command
    (`--attn-name eva --num-landmarks 49 --adaptive-proj default --window-size 7 --attn-2d --use-rpe`, README.md:117)."""
    if attn_name == 'eva':
        spec = dict(adaptive_proj='default', num_landmarks=49, use_t5_rpe=False, use_rpe=True, window_size=7, attn_2d=True,
                    overlap_window=False, fp32=False)
    elif attn_name == 'lara':
        spec = dict(num_landmarks=49, kernel_size=None, proposal_gen='pool-mixed', use_antithetics=False, use_multisample=False,
                    pool_module_type='light', mis_type='mis-opt', alpha_coeff=2.0, fp32=False)
    elif attn_name == 'performer':     # the defaults of kernelized_attention.py:322-330
        spec = dict(fp32=False, approx_attn_dim=64, proj_method='favorp', cos_weighting=False, sample_scheme='default')
    elif attn_name == 'scatterbrain':  # scatterbrain_attention.py:166-180 with the window flags of the DeiT + EVA command
        spec = dict(fp32=False, use_rpe=True, window_size=7, attn_2d=True, overlap_window=False, approx_attn_dim=64,
                    proj_method='favorp', cos_weighting=False, sample_scheme='default')
    else:
        spec = dict(fp32=False)
    spec.update(attn_kw)
    return Namespace(num_classes=num_classes, input_size=input_size, patchify_stem='default', no_pos_emb=False, drop_rate=0.0,
                     attn_drop_rate=0.0, drop_path_rate=0.1, use_glu=False, num_heads=None, attn_name=attn_name,
                     attn_specific_args=Namespace(**spec))
