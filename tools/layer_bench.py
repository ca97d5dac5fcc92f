"""Module-level forward timings of the other BASELINE configurations (development / DESIGN.md numbers; bench.py is the
contract and measures c3).  For each config: tokens/s of `module.forward(x)` on the device (x resident in HBM, fp16) and
the fraction of the module-level HBM roofline of SURVEY.md section 8(d): unfused x -> qkv -> o -> y traffic = 10 C s bytes
per token.

    python tools/layer_bench.py [iters]
"""
import argparse
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200'))
sys.path.insert(0, ROOT)
import efficient_attention as ea  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('iters', type=int, nargs='?', default=20)
args = ap.parse_args()
dev = torch.device('cuda', 0)
peak = 6469.3e9
try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] * 1e9
except Exception:
    pass


def reinit(m):
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() == 2 and 'bias' not in name:
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / p.shape[1] ** 0.5))
    return m


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def report(name, tokens, C, sec, note=''):
    tps = tokens / sec
    frac = tps * 10 * C * 2 / peak
    print(json.dumps({'config': name, 'tokens_per_s': tps, 'ms': sec * 1e3, 'module_bytes_per_token': 10 * C * 2,
                      'frac_of_module_hbm_roofline': frac, 'note': note}), flush=True)


with torch.no_grad(), warnings.catch_warnings():
    warnings.simplefilter('ignore')
    # c2: DeiT-tiny-p16 EVA, N = 196 (14 x 14), fused path (chunk 2)
    m = reinit(ea.AttentionFactory.build_attention('eva', dict(
        dim=192, num_heads=3, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False, use_rpe=True, window_size=7, attn_2d=True,
        overlap_window=False, adaptive_proj='default', num_landmarks=49, use_t5_rpe=False))).to(dev).half().eval()
    x = torch.randn(2048, 14, 14, 192, device=dev, dtype=torch.float16)
    report('c2 EVA N=196 C=192 B=2048 (fused core)', 2048 * 196, 192, timed(lambda: m(x), args.iters))
    from efficient_attention import _abi
    q, k, v, _ = m._qkv_heads(x.reshape(2048, 196, 192))
    geom = _abi.eva_geometry(q, seq_shape=(14, 14), window=7, ext=0, chunk=2, chunk_ext=0)
    ada, bias = m._adaptive(), m._local_bias().float().contiguous()
    sec = timed(lambda: _abi.eva_forward(q, k, v, geom, ada, bias=bias), args.iters)
    print(json.dumps({'config': 'c2 core only (eva_forward, q/k/v resident)', 'tokens_per_s': 2048 * 196 / sec, 'ms': sec * 1e3,
                      'frac_of_core_hbm_roofline': 2048 * 196 / sec * 1536 / peak}), flush=True)
    del q, k, v
    # c3 at module level, for comparison with the core number of bench.py
    x = torch.randn(1024, 28, 28, 192, device=dev, dtype=torch.float16)
    report('c3 EVA N=784 C=192 B=1024 (fused core)', 1024 * 784, 192, timed(lambda: m(x), args.iters))
    del m, x
    # c4: DeiT-small-p16 LARA, N = 196, C = 384, 6 heads, mis-opt, pool-mixed proposals
    m = reinit(ea.AttentionFactory.build_attention('lara', dict(
        dim=384, num_heads=6, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False, num_landmarks=49, proposal_gen='pool-mixed',
        use_antithetics=False, use_multisample=False, pool_module_type='light', mis_type='mis-opt', alpha_coeff=1.0))).to(dev).half().eval()
    x = torch.randn(512, 14, 14, 384, device=dev, dtype=torch.float16)
    report('c4 LARA N=196 C=384 B=512 (fused tcgen05 LARA kernel)', 512 * 196, 384, timed(lambda: m(x), max(3, args.iters // 4)))
    del m, x
    # c5: causal EVA LM layer, T = 4096, C = 512, 8 heads, chunk 256, window 256
    ns = argparse.Namespace(adaptive_proj='qk', num_chunks=None, chunk_size=256, causal=True, use_t5_rpe=False, window_size=256,
                            overlap_window=False)
    m = reinit(ea.CausalEVAttention(embed_dim=512, num_heads=8, dropout=0.0, self_attention=True, attn_args=ns)).to(dev).half().eval()
    x = torch.randn(4096, 16, 512, device=dev, dtype=torch.float16)
    report('c5 causal EVA T=4096 C=512 B=16 (CTA-per-chunk statistics + tcgen05 window kernel)', 16 * 4096, 512,
           timed(lambda: m(x, x, x, need_weights=False)[0], max(3, args.iters // 4)))
