"""c4 (LARA, N=196, C=384, 6 heads, 49 landmarks): run a few forwards (for an ncu launch list)."""
import os, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200')); sys.path.insert(0, ROOT)
import efficient_attention as ea
dev = torch.device('cuda', 0)
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    m = ea.AttentionFactory.build_attention('lara', dict(dim=384, num_heads=6, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False,
        num_landmarks=49, proposal_gen='pool-mixed', use_antithetics=False, use_multisample=False, pool_module_type='light',
        mis_type='mis-opt', alpha_coeff=1.0)).to(dev).half().eval()
x = torch.randn(512, 14, 14, 384, device=dev, dtype=torch.float16)
with torch.no_grad():
    for _ in range(3): m(x)
torch.cuda.synchronize()
