"""Shared pytest plumbing: path setup, the `gpu` marker, golden-fixture loading."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'efficient-attention_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def golden_names(prefixes=None):
    names = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz'))
    if prefixes:
        names = [n for n in names if n.startswith(tuple(prefixes))]
    return names


def load_golden(name, dtype=torch.float64):
    """-> (cfg dict, state_dict of tensors, dict(x, mask, noise, y))."""
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    cfg = json.loads(bytes(z['cfg']).decode())
    sd, arr = {}, {'mask': None, 'noise': None}
    for k in z.files:
        if k == 'cfg':
            continue
        t = torch.from_numpy(z[k])
        if k.startswith('sd::'):
            sd[k[4:]] = t.to(dtype) if t.is_floating_point() else t
        else:
            arr[k] = t.to(dtype) if t.is_floating_point() else t
    return cfg, sd, arr


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))
