"""The UNMODIFIED reference (oracle/_ref, pure PyTorch) run ON THE GPU beside this package: c3 EVA layer forward and the DeiT-tiny-p8
model forward / training step, fp16 autocast (development tool; the graded reference arm is the CPU one in bench.py).

    python tools/ref_gpu_compare.py            # runs both arms (each in its own process: the two packages share a name)
"""
import json
import os
import statistics
import subprocess
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, n, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs)


def causal_arm(which, ea, dev):
    """BASELINE config c5: one CausalEVAttention layer (embed 512, 8 heads, window = chunk = 256, T5 bias), T = 4096, batch 16, fp16
    autocast: forward (eval) and forward + backward (train)."""
    import warnings
    from argparse import Namespace
    import bench
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        torch.manual_seed(0)
        m = bench.lively_init(ea.CausalEVAttention(512, 8, self_attention=True, attn_args=Namespace(
            adaptive_proj='qk', num_chunks=None, chunk_size=256, causal=True, use_t5_rpe=True, window_size=256,
            overlap_window=False))).to(dev)
    x = torch.randn(4096, 16, 512, device=dev)
    out = {'arm': which, 'attn': 'causal', 'package': os.path.dirname(ea.__file__)}
    m.eval()

    def fwd():
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            return m(x, x, x, need_weights=False)
    out['layer_fwd_ms'] = timed(fwd, 20)
    m.train()

    def fwd_bwd():
        xx = x.detach().requires_grad_(True)
        with torch.autocast('cuda', dtype=torch.float16):
            y = m(xx, xx, xx, need_weights=False)[0]
        y.float().pow(2).mean().backward()
    try:
        out['layer_fwd_bwd_ms'] = timed(fwd_bwd, 10, warm=3)
        out['peak_gib'] = torch.cuda.max_memory_allocated() / 2 ** 30
    except RuntimeError as e:
        out['layer_fwd_bwd_ms'] = f'failed: {str(e)[:80]}'
    print(json.dumps(out))


def arm(which, attn='eva'):
    import bench
    from oracle import ref_loader
    dev = torch.device('cuda', 0)
    if which == 'reference':
        ea = ref_loader.reference_attention()
    else:
        bench.use_product_package()
        import efficient_attention as ea
    out = {'arm': which, 'attn': attn, 'package': os.path.dirname(ea.__file__)}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        vm = ref_loader.vit_models()
        torch.manual_seed(0)
        if attn == 'eva':
            layer = bench.lively_init(ea.AttentionFactory.build_attention('eva', dict(bench.EVA_ARGS))).to(dev).eval()
            model = vm.evit_tiny_p8(ref_loader.deit_args('eva')).to(dev)
        elif attn == 'causal':
            layer = model = None
        elif attn == 'softmax':                   # the dense baseline of the reference's ViT (abstract_attention.py), N = 784
            layer = None
            model = vm.evit_tiny_p8(ref_loader.deit_args('softmax')).to(dev)
        elif attn in ('performer', 'scatterbrain', 'ra'):   # the random-feature baselines of the registry in DeiT-tiny-p8
            layer = None
            model = vm.evit_tiny_p8(ref_loader.deit_args(attn)).to(dev)
        else:                                     # LARA: BASELINE config c4 = DeiT-small-p16 (196 tokens, 384 channels, 6 heads)
            layer = None
            model = vm.evit_small_p16(ref_loader.deit_args('lara')).to(dev)
        out['model'] = {'eva': 'evit_tiny_p8', 'lara': 'evit_small_p16', 'softmax': 'evit_tiny_p8', 'performer': 'evit_tiny_p8', 'scatterbrain': 'evit_tiny_p8', 'ra': 'evit_tiny_p8'}.get(attn, 'CausalEVAttention layer')
    for B in ((128, 1024) if layer is not None else ()):
        x = torch.randn(B, 28, 28, 192, device=dev)

        def layer_fwd():
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
                return layer(x)
        try:
            ms = timed(layer_fwd, 20)
            out[f'layer_fwd_ms_B{B}'] = ms
            out[f'layer_tokens_per_s_B{B}'] = B * 784 / (ms * 1e-3)
        except RuntimeError as e:
            out[f'layer_fwd_ms_B{B}'] = f'failed: {str(e)[:80]}'
        del x
        torch.cuda.empty_cache()
    if attn == 'causal':
        return causal_arm(which, ea, dev)
    img = torch.randn(128, 3, 224, 224, device=dev)
    model.eval()

    def model_fwd():
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            return model(img)
    ms = timed(model_fwd, 20)
    out.update(model_fwd_ms_B128=ms, model_images_per_s=128 / (ms * 1e-3))
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    y = torch.randint(0, 1000, (128,), device=dev)

    def train_step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.float16):
            loss = torch.nn.functional.cross_entropy(model(img).float(), y)
        loss.backward()
        opt.step()
    try:
        ms = timed(train_step, 10, warm=3)
        out.update(train_step_ms_B128=ms, train_images_per_s=128 / (ms * 1e-3))
    except RuntimeError as e:
        out['train_step_ms_B128'] = f'failed: {str(e)[:80]}'
    print(json.dumps(out))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] in ('reference', 'ours'):
        arm(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'eva')
    else:
        attn = sys.argv[1] if len(sys.argv) > 1 else 'eva'
        for which in ('reference', 'ours'):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), which, attn], capture_output=True, text=True, timeout=1200)
            lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
            print(lines[-1] if lines else f'{which}: failed\n{r.stderr[-1500:]}')
