"""Parity + timing of the LARA module (c4) with the tcgen05 core against the generic CUDA-core kernels (development tool)."""
import os, subprocess, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200')); sys.path.insert(0, ROOT)
import efficient_attention as ea
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mode = sys.argv[2] if len(sys.argv) > 2 else 'main'
dev = torch.device('cuda', 0)
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    m = ea.AttentionFactory.build_attention('lara', dict(dim=384, num_heads=6, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False,
        num_landmarks=49, proposal_gen='pool-mixed', use_antithetics=False, use_multisample=False, pool_module_type='light',
        mis_type='mis-opt', alpha_coeff=1.0))
g = torch.Generator().manual_seed(0)
with torch.no_grad():
    for name, p in m.named_parameters():
        if p.dim() == 2:
            p.copy_(torch.randn(p.shape, generator=g) * (1.0 / p.shape[1] ** 0.5))
m = m.to(dev).half().eval()
PB = 4
with torch.no_grad():
    torch.manual_seed(1)
    xs = torch.randn(PB, 14, 14, 384, device=dev, dtype=torch.float16)
    out = m(xs)
    torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    if mode == 'ref':
        torch.save(out.float().cpu(), os.path.join(ROOT, 'gpurun_out', 'lara_ref.pt'))
        sys.exit(0)
    subprocess.run([sys.executable, __file__, str(B), 'ref'], env=dict(os.environ, EVA_SM100_DISABLE_FUSED='1'), check=True)
    ref = torch.load(os.path.join(ROOT, 'gpurun_out', 'lara_ref.pt'))
    got = out.float().cpu()
    err = float((got - ref).norm() / ref.norm())
    print(f'parity (B={PB}) tcgen05 core vs generic, module output: rel-L2 {err:.3e}, max abs {float((got - ref).abs().max()):.3e}, nan {int(torch.isnan(got).sum())}', flush=True)
    x = torch.randn(B, 14, 14, 384, device=dev, dtype=torch.float16)
    for _ in range(3): m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): m(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'module forward B={B}: {ms:.3f} ms  ({B * 196 / ms / 1e3:.1f} M tokens/s)', flush=True)
    from efficient_attention import _abi
    q, k, v, _ = m._qkv_heads(x.reshape(B, 196, 384))
    core = lambda: _abi.lara_forward(q, k, v, seq_shape=(14, 14), landmarks=49, per_token_proj=False, mixed=1, mis_type='mis-opt',
                                     sample_mode=_abi.LARA_SAMPLE_SINGLE, zero_padded=False, alpha_coeff=1.0, proj=m._proj_params(True),
                                     pad_mask=None, noise=None)
    for _ in range(3): core()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): core()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f'core only (lara_forward, q/k/v resident) B={B}: {ms:.3f} ms  ({B * 196 / ms / 1e3:.1f} M tokens/s, '
          f'{B * 196 * 3072 / (ms * 1e-3) / 6469.3e9 * 100:.1f} % of the core HBM roofline)', flush=True)
    assert err < 3e-3, err
