"""GPU parity tests (`-m gpu`) of the random-feature cores (csrc/rfa_kernels.cu) through the C ABI, against the float64 oracle
(oracle/rfa_oracle.py) on identical, pre-quantised q / k / v, plus gradients of the modules against autograd through the oracle.

Tolerances (relative L2, oracle in float64): float32 I/O 2e-5; float16 I/O 1e-3 (north_star); bfloat16 I/O 6e-3 (output rounding).
"""
import math

import pytest
import torch

from conftest import load_golden, rel_l2
from helpers import build_module, set_draws
from oracle import eva_oracle as O
from oracle import rfa_oracle as R

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 2e-5, torch.float16: 1e-3, torch.bfloat16: 6e-3}


def _dev():
    return torch.device('cuda', 0)


def _qkv(B, N, H, D, dtype, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    packed = (scale * torch.randn(B, N, 3, H, D, generator=g)).to(dtype)
    dev = packed.to(_dev())
    ref = [packed[:, :, i].double().transpose(1, 2) for i in range(3)]          # [B, H, N, D] float64 of the SAME rounded values
    return (dev[:, :, 0], dev[:, :, 1], dev[:, :, 2]), ref


def _heads(o, B, N, H, D):
    return o.transpose(1, 2).reshape(B, N, H * D)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('method,m,cos', [('favorp', 64, False), ('favorp', 48, True), ('relu', 64, False), ('fourier', 32, False),
                                          ('dpfp', 256, False), ('relu-only', 0, True), ('sigmoid-only', 0, False)])
def test_performer_core_vs_oracle(dtype, method, m, cos):
    from efficient_attention import _abi
    B, N, H, D = 3, 203, 2, 64
    (q, k, v), (qr, kr, vr) = _qkv(B, N, H, D, dtype, 5)
    g = torch.Generator().manual_seed(7)
    proj = torch.randn(H, m, D, generator=g) if method in ('favorp', 'relu', 'fourier') else None
    mask = torch.zeros(B, N, dtype=torch.bool)
    mask[1, -17:] = True
    nu = 2
    out = _abi.rfa_forward(q, k, v, method=method, proj=None if proj is None else proj.to(_dev()), nu=nu, cos_weighting=cos,
                           pad_mask=mask.to(_dev()))
    ref = R.performer_core(qr, kr, vr, method=method, proj=None if proj is None else proj.double(), nu=nu, cos_weighting=cos,
                           pad_mask=mask)
    err = rel_l2(out.cpu(), _heads(ref, B, N, H, D))
    assert err < TOL[dtype], (method, dtype, err)


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('B,N,masked', [(2, 784, False), (3, 203, True), (1, 128, False), (120, 196, True)])
def test_performer_tcgen05_path_vs_oracle(dtype, B, N, masked):
    """'favorp', 64 features, head_dim 64, 16-bit: the tcgen05 kernel (launch counter), incl. ragged tiles, padded keys and more
    items than resident CTAs (B = 120, h = 3: every CTA loops), against the float64 oracle on the same rounded q / k / v."""
    from efficient_attention import _abi
    H, D = 3, 64
    (q, k, v), (qr, kr, vr) = _qkv(B, N, H, D, dtype, 21)
    proj = torch.randn(H, 64, D, generator=torch.Generator().manual_seed(5))
    mask = None
    if masked:
        mask = torch.zeros(B, N, dtype=torch.bool)
        mask[B - 1, -37:] = True
        mask[0, 5:9] = True
    before = _abi.rfa_tc_launches()
    out = _abi.rfa_forward(q, k, v, method='favorp', proj=proj.to(_dev()), pad_mask=None if mask is None else mask.to(_dev()))
    assert _abi.rfa_tc_launches() == before + 1
    worst = 0.0
    for b0 in range(0, B, 8):                      # the oracle in slices of the batch
        sl = slice(b0, min(B, b0 + 8))
        ref = R.performer_core(qr[sl], kr[sl], vr[sl], method='favorp', proj=proj.double(), pad_mask=None if mask is None else mask[sl])
        worst = max(worst, rel_l2(out[sl].cpu(), _heads(ref, ref.shape[0], N, H, D)))
    assert worst < TOL[dtype], (dtype, B, N, worst)


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
def test_performer_tcgen05_running_stabiliser_rescales(dtype):
    """Keys whose magnitude grows along the sequence: the running maximum of the tcgen05 kernel rises at every 128-token tile, so the
    accumulators in tensor memory are rescaled six times per item; padded keys (which still count for the stabiliser) carry the
    largest rows.  Against the float64 oracle."""
    from efficient_attention import _abi
    B, N, H, D = 3, 784, 3, 64
    g = torch.Generator().manual_seed(41)
    packed = torch.randn(B, N, 3, H, D, generator=g)
    packed[:, :, 1] *= (0.3 + 1.2 * torch.arange(N).float() / N).view(1, N, 1, 1)
    packed = packed.to(dtype)
    dev = packed.to(_dev())
    q, k, v = dev[:, :, 0], dev[:, :, 1], dev[:, :, 2]
    qr, kr, vr = (packed[:, :, i].double().transpose(1, 2) for i in range(3))
    proj = torch.randn(H, 64, D, generator=torch.Generator().manual_seed(6))
    mask = torch.zeros(B, N, dtype=torch.bool)
    mask[1, -100:] = True
    before = _abi.rfa_tc_launches()
    out = _abi.rfa_forward(q, k, v, method='favorp', proj=proj.to(_dev()), pad_mask=mask.to(_dev()))
    assert _abi.rfa_tc_launches() == before + 1
    ref = R.performer_core(qr, kr, vr, method='favorp', proj=proj.double(), pad_mask=mask)
    err = rel_l2(out.cpu(), _heads(ref, B, N, H, D))
    assert err < TOL[dtype], (dtype, err)


@pytest.mark.parametrize('D,m', [(16, 24), (32, 64), (128, 128)])
def test_performer_core_other_head_dims(D, m):
    from efficient_attention import _abi
    B, N, H = 2, 77, 3
    (q, k, v), (qr, kr, vr) = _qkv(B, N, H, D, torch.float32, 9)
    proj = torch.randn(H, m, D, generator=torch.Generator().manual_seed(1))
    out = _abi.rfa_forward(q, k, v, method='favorp', proj=proj.to(_dev()))
    ref = R.performer_core(qr, kr, vr, method='favorp', proj=proj.double())
    assert rel_l2(out.cpu(), _heads(ref, B, N, H, D)) < 2e-5


def test_performer_c3_shape_linearity_and_oracle_slice():
    """BASELINE c3 shape (N = 784, h = 3, d = 64), batch 64, fp16: a slice of the batch against the oracle, and the size-independent
    property out(v1 + v2) = out(v1) + out(v2) on the whole batch (the map is linear in v for fixed q, k)."""
    from efficient_attention import _abi
    B, N, H, D = 64, 784, 3, 64
    (q, k, v), (qr, kr, vr) = _qkv(B, N, H, D, torch.float16, 11)
    proj = torch.randn(H, 64, D, generator=torch.Generator().manual_seed(2))
    out = _abi.rfa_forward(q, k, v, method='favorp', proj=proj.to(_dev()))
    ref = R.performer_core(qr[:2], kr[:2], vr[:2], method='favorp', proj=proj.double())
    assert rel_l2(out[:2].cpu(), _heads(ref, 2, N, H, D)) < 1e-3
    v2 = torch.randn_like(v)
    o2 = _abi.rfa_forward(q, k, v2, method='favorp', proj=proj.to(_dev()))
    o12 = _abi.rfa_forward(q, k, (v.float() + v2.float()).half(), method='favorp', proj=proj.to(_dev()))
    assert rel_l2(o12.float().cpu(), (out.float() + o2.float()).cpu()) < 2e-3


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('mode', ['mean', 'given', 'gather'])
def test_ra_core_vs_oracle(dtype, mode):
    from efficient_attention import _abi
    B, N, H, D = 2, 150, 3, 64
    (q, k, v), (qr, kr, vr) = _qkv(B, N, H, D, dtype, 13, scale=0.7)
    g = torch.Generator().manual_seed(3)
    noise = 0.5 * torch.randn(B, H, N, D, generator=g)
    k_ind = torch.randint(0, N, (B, H, N), generator=g)
    tc_before = _abi.load().eva_debug_window_tc_count()
    if mode == 'given':
        extra = torch.randn(B, N, H * D, generator=g).to(dtype)
        ex_ref = extra.double().view(B, N, H, D).transpose(1, 2)
        out = _abi.ra_forward(q, k, v, mode='given', extra=extra.to(_dev()), noise=noise.to(_dev()))
        w_extra = ex_ref
    elif mode == 'mean':
        out = _abi.ra_forward(q, k, v, mode='mean', noise=noise.to(_dev()))
        w_extra = kr.mean(-2, keepdim=True)
    else:
        out = _abi.ra_forward(q, k, v, mode='gather', k_ind=k_ind.to(_dev()), noise=noise.to(_dev()))
        w_extra = torch.gather(kr, 2, k_ind.unsqueeze(-1).expand(-1, -1, -1, D))
    if dtype != torch.float32:       # head_dim 64, 16-bit: the dense tcgen05 window kernel with a per-key addend did the softmax pass
        assert _abi.load().eva_debug_window_tc_count() > tc_before
    s = D ** -0.5
    w = qr + w_extra + noise.double()
    ref = torch.softmax(s * (w @ kr.transpose(-1, -2)) - 0.5 * s * (kr * kr).sum(-1).unsqueeze(-2), -1) @ vr
    err = rel_l2(out.cpu(), _heads(ref, B, N, H, D))
    assert err < TOL[dtype], (mode, dtype, err)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('geom', ['2d_w7', '1d_w16_mask', '2d_w8_d32', '1d_w64_odd', '1d_w16_mask_m64', '2d_w7_many'])
def test_scatterbrain_core_vs_oracle(dtype, geom):
    from efficient_attention import _abi
    if geom == '2d_w7':
        B, H, D, shape, w, m = 2, 3, 64, (28, 28), 7, 64
    elif geom == '1d_w16_mask':
        B, H, D, shape, w, m = 3, 2, 64, (160,), 16, 48
    elif geom == '1d_w64_odd':             # 64-token windows, an odd number of them (the last pair is half empty)
        B, H, D, shape, w, m = 2, 2, 64, (320,), 64, 64
    elif geom == '1d_w16_mask_m64':
        B, H, D, shape, w, m = 3, 2, 64, (176,), 16, 64
    elif geom == '2d_w7_many':             # more items than resident CTAs: every CTA of the tcgen05 kernels loops
        B, H, D, shape, w, m = 110, 3, 64, (14, 14), 7, 64
    else:
        B, H, D, shape, w, m = 2, 2, 32, (16, 16), 8, 32
    N = math.prod(shape)
    if geom == '2d_w7_many' and dtype == torch.float32:
        pytest.skip('the many-items case is about the tcgen05 kernels')
    (q, k, v), (qr, kr, vr) = _qkv(B, N, H, D, dtype, 17, scale=0.8)
    g = torch.Generator().manual_seed(4)
    proj = torch.randn(H, m, D, generator=g)
    L = w * w if len(shape) == 2 else w
    bias = 0.5 * torch.randn(H, L, L, generator=g)
    mask = None
    if 'mask' in geom:
        mask = torch.zeros(B, N, dtype=torch.bool)
        mask[1, -21:] = True
        mask[2, 40:60] = True
    tc_before = _abi.load().eva_debug_sb_tc_launches()
    out = _abi.scatterbrain_forward(q, k, v, seq_shape=shape, window=w, proj=proj.to(_dev()), bias=bias.to(_dev()),
                                    pad_mask=None if mask is None else mask.to(_dev()))
    if dtype != torch.float32 and D == 64 and m == 64:
        assert _abi.load().eva_debug_sb_tc_launches() == tc_before + 1      # the tcgen05 window kernel ran
    worst = 0.0
    for b0 in range(0, B, 16):
        sl = slice(b0, min(B, b0 + 16))
        ref = R.scatterbrain_core(qr[sl], kr[sl], vr[sl], proj=proj.double(), seq_shape=shape, window=w, ext=0,
                                  pad_mask=None if mask is None else mask[sl], bias=bias.double())
        worst = max(worst, rel_l2(out[sl].cpu(), _heads(ref, ref.shape[0], N, H, D)))
    assert worst < TOL[dtype], (geom, dtype, worst)


@pytest.mark.parametrize('name', ['perf_favorp_mask', 'perf_fourier_learn', 'perf_mlp_fourier', 'perf_relu_only_cos', 'ra_mean',
                                  'ra_expect_2d', 'ra_sample_train', 'sb_2d_rpe', 'sb_1d_pad_mask'])
def test_module_gradients_match_autograd_through_the_oracle(name):
    """d loss / d x and d loss / d (every parameter) of the drop-in module (forward: CUDA kernels; backward: the module's float32
    restatement) against autograd through the float64 oracle, loss = sum(y * fixed random tensor)."""
    cfg, sd, a = load_golden(name, dtype=torch.float32)
    m = build_module(cfg)
    m.load_state_dict(sd)
    m = m.to(_dev())
    m.train()
    set_draws(m, cfg, a, _dev())
    if cfg['kind'] in ('performer', 'scatterbrain') and a.get('proj') is None and 'eval_proj' in sd:
        m._proj_override = sd['eval_proj'].to(_dev())         # training mode would draw a fresh projection
    if cfg['kind'] == 'ra' and a.get('noise') is None:
        m._draw_override = (a['k_ind'].to(_dev()) if a.get('k_ind') is not None else None, None)
    x = a['x'].to(_dev()).requires_grad_(True)
    mask = a['mask'].to(_dev()) if a['mask'] is not None else None
    y = m(x, mask)
    gsel = torch.randn(y.shape, generator=torch.Generator().manual_seed(1)).to(_dev())
    (y * gsel).sum().backward()
    # oracle side, float64
    sd64 = {k_: (v_.double().requires_grad_(True) if v_.is_floating_point() else v_) for k_, v_ in sd.items()}
    x64 = a['x'].double().requires_grad_(True)
    fwd = {'performer': lambda: R.performer_forward(sd64, cfg, x64, a['mask'], a.get('proj').double() if a.get('proj') is not None else None),
           'ra': lambda: R.ra_forward(sd64, cfg, x64, a['mask'], a.get('k_ind'), a['noise'].double() if a.get('noise') is not None else None),
           'scatterbrain': lambda: R.scatterbrain_forward(sd64, cfg, x64, a['mask'], a.get('proj').double() if a.get('proj') is not None else None)}
    y64 = fwd[cfg['kind']]()
    assert rel_l2(y.detach().cpu(), y64.detach()) < 5e-5
    (y64 * gsel.cpu().double()).sum().backward()
    assert rel_l2(x.grad.cpu(), x64.grad) < 2e-4, name
    for pname, p in m.named_parameters():
        ref_g = sd64[pname].grad
        if ref_g is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, pname
            continue
        assert p.grad is not None, pname
        assert rel_l2(p.grad.cpu(), ref_g) < 5e-4, (name, pname)


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('B,N', [(2, 203), (40, 384)])
def test_ra_sample_gumbel_max_vs_oracle(dtype, B, N):
    """`ra_sample` with EXPLICIT Gumbel noise: k_ind[n] = argmax_m (scale q_n . k_m + G_nm) against the float64 evaluation on the same
    rounded q / k.  Index work: exact, except where the two best candidates of a row are closer than the 16-bit MMA can resolve."""
    from efficient_attention import _abi
    H, D = 3, 64
    (q, k, _), (qr, kr, _) = _qkv(B, N, H, D, dtype, 31)
    u = torch.rand(B, H, N, N, generator=torch.Generator().manual_seed(9)).clamp(1e-7, 1 - 1e-7)
    gum = -torch.log(-torch.log(u))
    got = _abi.ra_sample(q, k, gumbel=gum.to(_dev())).cpu()
    val = D ** -0.5 * (qr @ kr.transpose(-1, -2)) + gum.double()
    want = val.argmax(-1)
    assert got.shape == want.shape and int(got.min()) >= 0 and int(got.max()) < N
    same = got == want
    assert float(same.float().mean()) > 0.999, float(same.float().mean())
    gap = val.max(-1).values - torch.gather(val, -1, got.unsqueeze(-1)).squeeze(-1)       # 0 where equal
    assert float(gap.max()) < (2e-2 if dtype == torch.float16 else 1e-1), float(gap.max())


def test_ra_sample_follows_pi():
    """The seeded draw: with every query equal (one row of pi, 24 keys) the empirical frequencies over 6144 x 8 independent draws
    match pi (total variation < 0.02; the expected TV of that many multinomial draws is ~0.01)."""
    from efficient_attention import _abi
    B, N, H, D = 8, 24, 1, 64
    g = torch.Generator().manual_seed(3)
    q1, k1 = 2.0 * torch.randn(D, generator=g), 1.5 * torch.randn(N, D, generator=g)
    reps = 256
    q = q1.view(1, 1, 1, D).expand(B, N, H, D).half().contiguous().to(_dev())
    k = k1.view(1, N, 1, D).expand(B, N, H, D).half().contiguous().to(_dev())
    pi = torch.softmax(D ** -0.5 * (q1.half().double() @ k1.half().double().t()), -1)
    counts = torch.zeros(N, dtype=torch.float64)
    for rep in range(reps):
        idx = _abi.ra_sample(q, k, seed=1000 + rep).cpu().reshape(-1)
        counts += torch.bincount(idx, minlength=N).double()
    freq = counts / counts.sum()
    tv = 0.5 * float((freq - pi).abs().sum())
    assert tv < 0.02, (tv, freq, pi)


def test_ra_default_sampling_draws_from_pi():
    """num_samples = 1 in eval mode without the test hook: the module draws one key per query from pi = softmax(scale q k^T)
    (library ops, as the reference's torch.multinomial).  With a peaked pi the draw is (almost surely) the arg-max key."""
    import efficient_attention as ea
    torch.manual_seed(0)
    m = ea.AttentionFactory.build_attention('ra', dict(dim=64, num_heads=1, num_samples=1)).to(_dev()).eval()
    x = torch.randn(2, 40, 64, device=_dev())
    with torch.no_grad():
        m.qkv.weight.mul_(60.0)
        y = m(x)
    assert torch.isfinite(y).all() and y.shape == x.shape
