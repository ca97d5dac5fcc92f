"""Bit-exact repeatability of the fused core, the module forward and the host pipeline at the bench size (development tool)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200')); sys.path.insert(0, ROOT)
import bench
from efficient_attention import _abi
from efficient_attention.streaming import HostPipeline
dev = torch.device('cuda', 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
layer = bench.build_layer(dev, torch.float16)
torch.manual_seed(1)
with torch.no_grad():
    x = torch.randn(B, 28, 28, 192, device=dev, dtype=torch.float16)
    q, k, v, _ = layer._qkv_heads(x.reshape(B, 784, 192))
    geom = _abi.eva_geometry(q, seq_shape=(28, 28), window=7, ext=0, chunk=4, chunk_ext=0)
    ada, bias = layer._adaptive(), layer._local_bias().float().contiguous()
    ref = _abi.eva_forward(q, k, v, geom, ada, bias=bias).clone()
    bad = 0
    for i in range(20):
        out = _abi.eva_forward(q, k, v, geom, ada, bias=bias)
        if not torch.equal(out, ref):
            d = (out != ref).view(B, -1).any(1).nonzero().flatten().tolist()
            print(f'core run {i}: differs in images {d[:10]} ({len(d)} images)', flush=True); bad += 1
    print(f'core: {bad} of 20 repeats differ', flush=True)
    # small-batch pieces of the same input must equal the corresponding rows of the big launch
    for lo in (0, 300, 900):
        sub = _abi.eva_forward(q[lo:lo + 64], k[lo:lo + 64], v[lo:lo + 64], _abi.eva_geometry(q[lo:lo + 64], seq_shape=(28, 28), window=7, ext=0, chunk=4, chunk_ext=0), ada, bias=bias)
        print(f'rows {lo}..{lo + 64} of the big launch equal a 64-image launch: {torch.equal(sub, ref[lo:lo + 64])}', flush=True)
    yref = layer(x).clone()
    bad = sum(0 if torch.equal(layer(x), yref) else 1 for _ in range(10))
    print(f'module: {bad} of 10 repeats differ', flush=True)
    x_host = x.cpu().pin_memory(); y_host = torch.empty_like(x_host).pin_memory()
    pipe = HostPipeline(layer, chunk=B // 8)
    pipe(x_host, y_host); torch.cuda.synchronize(); y1 = y_host.clone()
    bad = 0
    for _ in range(5):
        pipe(x_host, y_host); torch.cuda.synchronize()
        bad += 0 if torch.equal(y_host, y1) else 1
    print(f'pipeline: {bad} of 5 repeats differ; equals direct module forward: {torch.equal(y1, yref.cpu())}', flush=True)
    # a short-lived pipeline object (dropped while its copies are still in flight) must give the same bits
    for trial in range(3):
        HostPipeline(layer, chunk=B // 8)(x_host, y_host)
        torch.cuda.synchronize()
        neq = (y_host != y1)
        print(f'temporary pipeline {trial}: {int(neq.sum())} elements differ in images {neq.view(B, -1).any(1).nonzero().flatten().tolist()[:12]}, '
              f'max abs diff {float((y_host.float() - y1.float()).abs().max()):.3e}, nan {int(torch.isnan(y_host).sum())}', flush=True)
    for seed in range(6):
        torch.manual_seed(100 + seed)
        xs = torch.randn(B, 28, 28, 192).half().to(dev)
        ys = layer(xs)
        print(f'seed {100 + seed}: nan {int(torch.isnan(ys).sum())} inf {int(torch.isinf(ys).sum())} max |y| {float(ys.float().abs().max()):.3f}', flush=True)
