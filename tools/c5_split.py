"""c5 (causal EVA, T=4096, C=512, 8 heads, window = chunk = 256): time of the two generic kernels and of the projections."""
import argparse, os, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200')); sys.path.insert(0, ROOT)
import efficient_attention as ea
from efficient_attention import _abi
dev = torch.device('cuda', 0)
ns = argparse.Namespace(adaptive_proj='qk', num_chunks=None, chunk_size=256, causal=True, use_t5_rpe=False, window_size=256, overlap_window=False)
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    m = ea.CausalEVAttention(embed_dim=512, num_heads=8, dropout=0.0, self_attention=True, attn_args=ns).to(dev).half().eval()
B, N, H, D = 16, 4096, 8, 64
x = torch.randn(B, N, 512, device=dev, dtype=torch.float16)
def timed(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
with torch.no_grad():
    q = m.q_proj(x).view(B, N, H, D); k = m.k_proj(x).view(B, N, H, D); v = m.v_proj(x).view(B, N, H, D)
    geom = _abi.eva_geometry(q, seq_shape=(N,), window=256, ext=0, chunk=256, chunk_ext=0, causal=True, halo_left_only=True, mask_queries=True)
    ada = m._adaptive()
    print('projections (3 GEMMs): %.3f ms' % timed(lambda: (m.q_proj(x), m.k_proj(x), m.v_proj(x))))
    print('chunk_stats kernel:    %.3f ms' % timed(lambda: _abi.eva_chunk_stats(q, k, v, geom, ada)))
    kb, be = _abi.eva_chunk_stats(q, k, v, geom, ada)
    print('window_attn kernel:    %.3f ms' % timed(lambda: _abi.eva_window_attention(q, k, v, geom, k_bar=kb, beta=be)))
    print('eva_forward (both):    %.3f ms' % timed(lambda: _abi.eva_forward(q, k, v, geom, ada)))
    o = _abi.eva_forward(q, k, v, geom, ada)
    print('out_proj GEMM:         %.3f ms' % timed(lambda: m.out_proj(o)))
