"""A single GEMM shape for ncu: python tools/one_gemm.py M N K [f32out] [residual]"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
M, N, K = (int(x) for x in sys.argv[1:4])
f32 = len(sys.argv) > 4 and sys.argv[4] == "1"
res = len(sys.argv) > 5 and sys.argv[5] == "1"
a = torch.randn(M, K, device='cuda').bfloat16(); w = torch.randn(N, K, device='cuda').bfloat16()
bias = torch.randn(N, device='cuda')
out = torch.empty(M, N, device='cuda', dtype=torch.float32 if f32 else torch.bfloat16)
for _ in range(5):
    ops.gemm(a, w, out, bias=bias, residual=out if res else None)
torch.cuda.synchronize()
print("ok")
