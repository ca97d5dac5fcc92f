import os, sys, warnings
sys.path.insert(0, '/root/repo'); 
import torch, bench
from argparse import Namespace
bench.use_product_package()
import efficient_attention as ea
dev = torch.device('cuda', 0)
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    torch.manual_seed(0)
    m = bench.lively_init(ea.CausalEVAttention(512, 8, self_attention=True, attn_args=Namespace(adaptive_proj='qk', num_chunks=None, chunk_size=256, causal=True, use_t5_rpe=True, window_size=256, overlap_window=False))).to(dev).train()
x = torch.randn(4096, 16, 512, device=dev)
def step():
    xx = x.detach().requires_grad_(True)
    with torch.autocast('cuda', dtype=torch.float16):
        y = m(xx, xx, xx, need_weights=False)[0]
    y.float().pow(2).mean().backward()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = sorted(((e.device_time_total, e.count, e.key) for e in prof.key_averages() if e.device_time_total > 0), reverse=True)
print('total', sum(r[0] for r in rows)/1e3)
for t, n, k in rows[:16]: print(f'{t/1e3:8.3f} ms x{n:<3d} {k[:100]}')
