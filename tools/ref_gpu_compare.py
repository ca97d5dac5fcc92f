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


def arm(which):
    import bench
    from oracle import ref_loader
    dev = torch.device('cuda', 0)
    if which == 'reference':
        ea = ref_loader.reference_attention()
    else:
        bench.use_product_package()
        import efficient_attention as ea
    out = {'arm': which, 'package': os.path.dirname(ea.__file__)}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        torch.manual_seed(0)
        layer = bench.lively_init(ea.AttentionFactory.build_attention('eva', dict(bench.EVA_ARGS))).to(dev).eval()
        vm = ref_loader.vit_models()
        torch.manual_seed(0)
        model = vm.evit_tiny_p8(ref_loader.deit_args('eva')).to(dev)
    for B in (128, 1024):
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
    if len(sys.argv) > 1:
        arm(sys.argv[1])
    else:
        for which in ('reference', 'ours'):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), which], capture_output=True, text=True, timeout=1200)
            lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
            print(lines[-1] if lines else f'{which}: failed\n{r.stderr[-1500:]}')
