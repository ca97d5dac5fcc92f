"""Recipe for `oracle/_ref/`: the UNMODIFIED reference, made importable where timm / fairseq are absent.

TEST INFRASTRUCTURE, NOT PRODUCT.  Run in the build container (needs /root/reference):

    python oracle/make_ref.py

The reference is pure Python, so "building" it is a file copy plus a ~25-line `timm` shim:

    oracle/_ref/efficient_attention/   <- /root/reference/efficient-attention/efficient_attention/   (verbatim)
    oracle/_ref/models/                <- /root/reference/vit/models/                                (verbatim)
    oracle/_ref/timm/                  <- shim: trunc_normal_, DropPath, to_2tuple, register_model, _cfg
    oracle/_ref/MANIFEST.json          <- sha256 of every copied file + the reference commit, for the record

`oracle/_ref/` is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so it
travels to the GPU box like a built .so.  Consumers: `bench.py --impl reference` (the reference arm, `cpu_baseline.kind =
"reference"`), `bench.py`'s DeiT leg and `tests/test_model_gpu.py` (the reference's own ViT as the CALLER of the drop-in
package), `tests/test_ref_arm.py`.  Nothing under efficient-attention_b200/ imports it.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = '/root/reference'
OUT = os.path.join(HERE, '_ref')

TIMM_LAYERS = '''"""timm.models.layers shim: the three symbols the reference imports (abstract_attention.py:5, efficient_vit.py:11)."""
import collections.abc
from itertools import repeat

import torch
from torch import nn
from torch.nn.init import trunc_normal_  # noqa: F401


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return tuple(repeat(x, 2))


class DropPath(nn.Module):
    """Stochastic depth per sample (timm semantics: scale by 1 / keep_prob)."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask
'''

TIMM_REGISTRY = '''"""timm.models.registry shim: `register_model` records the constructor and returns it unchanged."""
_model_entrypoints = {}


def register_model(fn):
    _model_entrypoints[fn.__name__] = fn
    return fn


def model_entrypoint(name):
    return _model_entrypoints[name]
'''

TIMM_VIT = '''"""timm.models.vision_transformer shim: `_cfg` (pvt_legacy.py:8 only stores the dict)."""


def _cfg(url='', **kwargs):
    return dict(url=url, num_classes=1000, input_size=(3, 224, 224), **kwargs)
'''


def _sha(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'efficient-attention', 'efficient_attention'))


def build(verbose=True):
    """(Re)creates oracle/_ref from /root/reference.  No-op with a message when the reference is absent (GPU box)."""
    if not available():
        if verbose:
            print('oracle/make_ref.py: /root/reference not present; keeping the prebuilt oracle/_ref (if any)')
        return os.path.isdir(os.path.join(OUT, 'efficient_attention'))
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    manifest = {'files': {}}
    for src, dst in ((os.path.join(REF_ROOT, 'efficient-attention', 'efficient_attention'), 'efficient_attention'),
                     (os.path.join(REF_ROOT, 'vit', 'models'), 'models')):
        os.makedirs(os.path.join(OUT, dst))
        for f in sorted(os.listdir(src)):
            if f.endswith('.py'):
                shutil.copyfile(os.path.join(src, f), os.path.join(OUT, dst, f))
                manifest['files'][f'{dst}/{f}'] = _sha(os.path.join(src, f))
    os.makedirs(os.path.join(OUT, 'timm', 'models'))
    open(os.path.join(OUT, 'timm', '__init__.py'), 'w').close()
    open(os.path.join(OUT, 'timm', 'models', '__init__.py'), 'w').close()
    for name, text in (('layers.py', TIMM_LAYERS), ('registry.py', TIMM_REGISTRY), ('vision_transformer.py', TIMM_VIT)):
        with open(os.path.join(OUT, 'timm', 'models', name), 'w') as f:
            f.write(text)
    try:
        manifest['reference_commit'] = subprocess.run(['git', '-C', REF_ROOT, 'rev-parse', 'HEAD'], capture_output=True,
                                                      text=True, timeout=10).stdout.strip() or None
    except Exception:
        manifest['reference_commit'] = None
    manifest['note'] = 'verbatim copies; the only non-reference files are timm/* (shim written by oracle/make_ref.py)'
    with open(os.path.join(OUT, 'MANIFEST.json'), 'w') as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print(f'oracle/make_ref.py: wrote {OUT} ({len(manifest["files"])} reference files + timm shim)')
    return True


if __name__ == '__main__':
    ok = build()
    sys.exit(0 if ok else 1)
