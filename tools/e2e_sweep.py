"""e2e (host -> device -> host) throughput of EVA.forward through HostPipeline for several chunk counts / depths, plus the
raw pinned-copy bandwidth of the box in both directions at once (the bound)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200')); sys.path.insert(0, ROOT)
import bench
from efficient_attention.streaming import HostPipeline
dev = torch.device('cuda', 0)
B = 1024
layer = bench.build_layer(dev, torch.float16)
x_host = torch.randn(B, 28, 28, 192).half().pin_memory()
y_host = torch.empty_like(x_host).pin_memory()
# raw duplex copy bound
xd = torch.empty_like(x_host, device=dev); yd = torch.randn(B, 28, 28, 192, device=dev, dtype=torch.float16)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def duplex():
    with torch.cuda.stream(s1): xd.copy_(x_host, non_blocking=True)
    with torch.cuda.stream(s2): y_host.copy_(yd, non_blocking=True)
for _ in range(2): duplex()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): duplex()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f'raw duplex copy of one step: {dt * 1e3:.2f} ms  ({x_host.numel() * 2 / dt / 1e9:.1f} GB/s each way)  -> bound {B * 784 / dt / 1e6:.1f} M tokens/s', flush=True)
x2 = x_host.clone().pin_memory(); y2 = torch.empty_like(x_host).pin_memory()
for n_chunks, depth, defer, two in ((8, 2, False, False), (8, 2, True, False), (8, 2, True, True), (8, 3, True, True), (16, 2, False, False), (16, 3, True, True), (4, 2, False, False), (4, 2, True, True)):
    pipe = HostPipeline(layer, chunk=B // n_chunks, depth=depth, defer_join=defer)
    xs, ys = ([x_host, x2], [y_host, y2]) if two else ([x_host, x_host], [y_host, y_host])
    for i in range(3): pipe(xs[i & 1], ys[i & 1])
    pipe.join(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(8): pipe(xs[i & 1], ys[i & 1])
    pipe.join(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 8
    print(f'chunks {n_chunks:2d} depth {depth} defer_join {defer} two buffer pairs {two}: {dt * 1e3:.2f} ms/step  {B * 784 / dt / 1e6:.1f} M tokens/s', flush=True)
for n_chunks, depth, graphs in ((8, 2, False),):
    if True:
        pipe = HostPipeline(layer, chunk=B // n_chunks, depth=depth)
        for _ in range(3): pipe(x_host, y_host)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(8): pipe(x_host, y_host)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 8
        y1 = y_host.clone()
        HostPipeline(layer, chunk=B // 8)(x_host, y_host)
        torch.cuda.synchronize()
        same = bool((y1 == y_host).all())
        with torch.no_grad():
            want = layer(x_host.to(dev)).cpu()
        neq = (y1 != y_host)
        print(f'   differing elements {int(neq.sum())} in images {neq.view(B, -1).any(1).nonzero().flatten().tolist()[:10]}; '
              f'loop result == direct forward: {torch.equal(y1, want)}; fresh pipeline == direct forward: {torch.equal(y_host, want)}; '
              f'max |diff| {float((y1.float() - y_host.float()).abs().max()):.3e}', flush=True)
        print(f'chunks {n_chunks:2d} depth {depth} graphs {graphs}: {dt * 1e3:.2f} ms/step  {B * 784 / dt / 1e6:.1f} M tokens/s  identical to eager: {same}', flush=True)
