"""Parity + timing of the tcgen05 causal window kernel against the generic CUDA-core path (development tool).
    python tools/causal_bench.py [batch] [chunk]"""
import argparse, os, subprocess, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200')); sys.path.insert(0, ROOT)
import efficient_attention as ea
from efficient_attention import _abi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mode = sys.argv[3] if len(sys.argv) > 3 else 'main'
dev = torch.device('cuda', 0)
ns = argparse.Namespace(adaptive_proj='qk', num_chunks=None, chunk_size=chunk, causal=True, use_t5_rpe=False, window_size=256, overlap_window=False)
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    m = ea.CausalEVAttention(embed_dim=512, num_heads=8, dropout=0.0, self_attention=True, attn_args=ns)
g = torch.Generator().manual_seed(0)
with torch.no_grad():
    for name, p in m.named_parameters():
        if p.dim() == 2:
            p.copy_(torch.randn(p.shape, generator=g) * (1.0 / p.shape[1] ** 0.5))
m = m.to(dev).half().eval()
N, H, D = 4096, 8, 64
PB = 2
def core(x):
    q = m.q_proj(x).view(x.shape[0], N, H, D); k = m.k_proj(x).view(x.shape[0], N, H, D); v = m.v_proj(x).view(x.shape[0], N, H, D)
    geom = _abi.eva_geometry(q, seq_shape=(N,), window=256, ext=0, chunk=chunk, chunk_ext=0, causal=True, halo_left_only=True, mask_queries=True)
    return q, k, v, geom
with torch.no_grad():
    torch.manual_seed(1)
    xs = torch.randn(PB, N, 512, device=dev, dtype=torch.float16)
    q, k, v, geom = core(xs)
    out, path = _abi.eva_forward(q, k, v, geom, m._adaptive(), return_path=True)
    torch.cuda.synchronize()
    if mode == 'ref':
        assert path == 0
        torch.save(out.float().cpu(), os.path.join(ROOT, 'gpurun_out', 'causal_ref.pt'))
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    subprocess.run([sys.executable, __file__, str(B), str(chunk), 'ref'], env=dict(os.environ, EVA_SM100_DISABLE_FUSED='1'), check=True)
    ref = torch.load(os.path.join(ROOT, 'gpurun_out', 'causal_ref.pt'))
    got = out.float().cpu()
    err = float((got - ref).norm() / ref.norm())
    worst = float((got - ref).abs().max())
    print(f'path {path}; parity (B={PB}, chunk {chunk}) tcgen05 vs generic: rel-L2 {err:.3e}, max abs {worst:.3e}, nan {int(torch.isnan(got).sum())}', flush=True)
    x = torch.randn(B, N, 512, device=dev, dtype=torch.float16)
    q, k, v, geom = core(x)
    ada = m._adaptive()
    for _ in range(3): _abi.eva_forward(q, k, v, geom, ada)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): _abi.eva_forward(q, k, v, geom, ada)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    kb, be = _abi.eva_chunk_stats(q, k, v, geom, ada)
    e0.record()
    for _ in range(20): _abi.eva_window_attention(q, k, v, geom, k_bar=kb, beta=be)
    e1.record(); torch.cuda.synchronize()
    ms_w = e0.elapsed_time(e1) / 20
    tok = B * N
    print(f'eva_forward {ms:.3f} ms ({tok / ms / 1e3:.1f} M tokens/s, {tok * 4096 / (ms * 1e-3) / 6469.3e9 * 100:.1f} % of the core HBM roofline); window kernel alone {ms_w:.3f} ms', flush=True)
    assert err < 3e-3, err
