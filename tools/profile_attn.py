"""Self-attention S=4096, d=40 (CFG batch 2, 8 heads) a few times: target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
DEV = "cuda:0"
S, d, H, B = int(os.environ.get("S", 4096)), int(os.environ.get("D", 40)), 8, 2
C = H * d
torch.manual_seed(0)
qkv = torch.randn(B * S, 3 * C, device=DEV).bfloat16()
out = torch.empty(B * S, C, device=DEV, dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, qkv, qkv, out, B, H, S, S, d, d ** -0.5, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, k_off=C, v_off=2 * C)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention(qkv, qkv, qkv, out, B, H, S, S, d, d ** -0.5, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, k_off=C, v_off=2 * C)
e1.record(); torch.cuda.synchronize()
print("avg us per attention (incl. V transpose):", e0.elapsed_time(e1) * 100)
