"""One self-attention backward (L0 shape) for ncu: python tools/one_attn_bwd.py [B] [S] [d]"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
d = int(sys.argv[3]) if len(sys.argv) > 3 else 40
H = 8; C = H * d
qkv = torch.randn(B * S, 3 * C, device='cuda').bfloat16()
do = torch.randn(B * S, C, device='cuda').bfloat16()
out = torch.empty(B * S, C, device='cuda', dtype=torch.bfloat16)
lse = torch.empty(B, H, S, device='cuda')
dqkv = torch.empty_like(qkv)
ops.attention_lse(qkv, qkv, qkv, out, lse, B, H, S, S, d, d ** -0.5, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, k_off=C, v_off=2 * C)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    e0.record()
    ops.attention_bwd(qkv, qkv, qkv, out, do, lse, dqkv, dqkv, dqkv, B, H, S, S, d, d ** -0.5, ldq=3 * C, ldk=3 * C, ldv=3 * C,
                      lddq=3 * C, lddk=3 * C, lddv=3 * C, k_off=C, v_off=2 * C, dk_off=C, dv_off=2 * C)
    e1.record(); torch.cuda.synchronize()
print(f"attention_bwd B{B} S{S} d{d}: {e0.elapsed_time(e1) * 1e3:.0f} us")
