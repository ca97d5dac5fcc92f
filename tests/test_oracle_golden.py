"""Pin the CPU oracle (oracle/eva_oracle.py, oracle/rfa_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import pytest
import torch

from conftest import golden_names, load_golden, rel_l2
from oracle import eva_oracle as O
from oracle import rfa_oracle as R

FORWARD = {
    'eva': lambda cfg, sd, a: O.eva_forward(sd, cfg, a['x'], a['mask'], a['noise']),
    'local': lambda cfg, sd, a: O.local_forward(sd, cfg, a['x'], a['mask']),
    'softmax': lambda cfg, sd, a: O.softmax_forward(sd, cfg, a['x'], a['mask']),
    'lara': lambda cfg, sd, a: O.lara_forward(sd, cfg, a['x'], a['mask'], a['noise']),
    'causal_eva': lambda cfg, sd, a: O.causal_eva_forward(sd, cfg, a['x'], a['mask'], a['noise']),
    'performer': lambda cfg, sd, a: R.performer_forward(sd, cfg, a['x'], a['mask'], a.get('proj'), f32_linear=True),
    'ra': lambda cfg, sd, a: R.ra_forward(sd, cfg, a['x'], a['mask'], a.get('k_ind'), a['noise']),
    'scatterbrain': lambda cfg, sd, a: R.scatterbrain_forward(sd, cfg, a['x'], a['mask'], a.get('proj')),
}


@pytest.mark.parametrize('name', golden_names())
def test_oracle_matches_reference_output(name):
    cfg, sd, a = load_golden(name)
    y = FORWARD[cfg['kind']](cfg, sd, a)
    assert y.shape == a['y'].shape
    # both sides are float64 evaluations of the same float32-representable inputs -- except 'performer', whose reference casts to
    # float32 for the linear-attention step whatever the module dtype (kernelized_attention.py:319): float32 rounding there
    assert rel_l2(y, a['y']) < (2e-6 if cfg['kind'] == 'performer' else 1e-11), name


def test_golden_set_covers_every_module_kind():
    kinds = {load_golden(n)[0]['kind'] for n in golden_names()}
    assert kinds == set(FORWARD)


def test_causal_consistency_property():
    """The reference's only self-check (causal_eva.py:916-950): position j of the full-sequence output
    equals position j of the output on any prefix longer than j."""
    cfg, sd, a = load_golden('causal_selfcheck')
    x = a['x']
    full = O.causal_eva_forward(sd, cfg, x)
    j = 25
    for t in (26, 40, 64, 65, 100):
        part = O.causal_eva_forward(sd, cfg, x[:t])
        assert torch.allclose(part[j], full[j], atol=1e-12, rtol=0), t
