"""Random-feature modules ('performer', 'scatterbrain', 'ra') on the c3 geometry (28 x 28 tokens, C = 192, h = 3), fp16 autocast:
this package against the UNMODIFIED reference package (oracle/_ref, its own PyTorch ops) on the same GPU, each in its own process,
plus the attention cores alone through the C ABI with q / k / v resident (development tool).

    python tools/rfa_bench.py [batch]
"""
import json
import os
import subprocess
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.ref_gpu_compare import timed  # noqa: E402

BASE = dict(dim=192, num_heads=3, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False)
SPECS = {
    'performer': dict(BASE, approx_attn_dim=64, proj_method='favorp'),
    'scatterbrain': dict(BASE, approx_attn_dim=64, window_size=7, attn_2d=True, use_rpe=True),
    'ra': dict(BASE, num_samples=-1),
    'ra_sample': dict(BASE, num_samples=1),          # the registry default: one key drawn from pi per query (torch.multinomial)
}


def arm(which, B):
    import bench
    from oracle import ref_loader
    dev = torch.device('cuda', 0)
    if which == 'reference':
        ea = ref_loader.reference_attention()
    else:
        bench.use_product_package()
        import efficient_attention as ea
    out = {'arm': which, 'batch': B}
    for name, spec in SPECS.items():
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            torch.manual_seed(0)
            layer = bench.lively_init(ea.AttentionFactory.build_attention(name.split('_')[0], dict(spec))).to(dev).eval()
        bb = B if not name.startswith('ra') else min(B, 128)
        x = torch.randn(bb, 28, 28, 192, device=dev)

        def fwd():
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
                return layer(x)
        try:
            ms = timed(fwd, 10, warm=3)
            out[name] = {'batch': bb, 'layer_fwd_ms': round(ms, 4), 'tokens_per_s': bb * 784 / (ms * 1e-3)}
        except RuntimeError as e:
            out[name] = {'batch': bb, 'failed': str(e)[:100]}
        if which == 'ours' and name != 'ra_sample':
            from efficient_attention import _abi
            xh = x.half()
            lh = layer.half()
            with torch.no_grad():
                q, k, v, _ = lh._qkv_heads(xh.reshape(bb, 784, 192))
                if name == 'performer':
                    core = lambda: _abi.rfa_forward(q, k, v, method='favorp', proj=layer.eval_proj.float())
                elif name == 'scatterbrain':
                    bias = lh._window_bias()
                    core = lambda: _abi.scatterbrain_forward(q, k, v, seq_shape=(28, 28), window=7, proj=layer.eval_proj.float(), bias=bias)
                else:
                    ex = _abi.eva_window_attention(q, k, k, _abi.eva_geometry(q, seq_shape=(784,), window=784, ext=0, chunk=0, chunk_ext=0, mask_is_neg_inf=True))
                    core = lambda: _abi.ra_forward(q, k, v, mode='given', extra=ex)
                ms = timed(core, 10, warm=3)
            out[name].update(core_ms=round(ms, 4), core_tokens_per_s=bb * 784 / (ms * 1e-3),
                             core_roofline_frac=bb * 784 * 1536 / (ms * 1e-3) / 6545.9e9)
        del x
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] in ('reference', 'ours'):
        arm(sys.argv[1], int(sys.argv[2]))
    else:
        B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
        for which in ('reference', 'ours'):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), which, str(B)], capture_output=True, text=True, timeout=1200)
            lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
            print(lines[-1] if lines else f'{which}: failed\n{r.stderr[-1500:]}')
