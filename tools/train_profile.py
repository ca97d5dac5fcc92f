"""Kernel-time breakdown of one DeiT-tiny-p8 + EVA training step (reference ViT, this package as `efficient_attention`), B = 128,
fp16 autocast: torch.profiler CUDA times grouped by kernel name (development tool).

    python tools/train_profile.py [batch]
"""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from oracle import ref_loader
    bench.use_product_package()
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    attn = sys.argv[2] if len(sys.argv) > 2 else 'eva'          # eva (evit_tiny_p8) | lara (evit_small_p16) | softmax (evit_tiny_p8)
    dev = torch.device('cuda', 0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        vm = ref_loader.vit_models()
        torch.manual_seed(0)
        model = (vm.evit_small_p16 if attn == 'lara' else vm.evit_tiny_p8)(ref_loader.deit_args(attn)).to(dev).train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    img = torch.randn(B, 3, 224, 224, device=dev)
    y = torch.randint(0, 1000, (B,), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.float16):
            loss = torch.nn.functional.cross_entropy(model(img).float(), y)
        loss.backward()
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
    rows.sort(key=lambda r: -r[1])
    total = sum(r[1] for r in rows)
    print(f'total device time {total / 1e3:.2f} ms over {sum(r[2] for r in rows)} kernels')
    for name, t, n in rows[:28]:
        print(f'{t / 1e3:8.3f} ms {100 * t / total:5.1f}%  x{n:<4d} {name[:110]}')


if __name__ == '__main__':
    main()
