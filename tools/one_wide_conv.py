"""The VAE decoder's widest conv3x3 shapes in isolation (image rows wider than a tile; CTA pairs): ncu target + timing.
python tools/one_wide_conv.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops, packing
DEV = "cuda:0"
torch.manual_seed(0)
for (B, H, W, Cin, Cout) in ((1, 512, 512, 128, 128), (1, 512, 512, 256, 256), (1, 256, 256, 512, 512)):
    x = torch.randn(B * H * W, Cin, device=DEV).bfloat16()
    w = packing.pack_conv3x3(torch.randn(Cout, Cin, 3, 3, device=DEV) * (9 * Cin) ** -0.5)
    bias = torch.randn(Cout, device=DEV)
    res = torch.randn(B * H * W, Cout, device=DEV)
    out = torch.empty(B * H * W, Cout, device=DEV)
    for _ in range(3):
        ops.gemm(x, w, out, bias=bias, residual=res, conv=(B, H, W))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(x, w, out, bias=bias, residual=res, conv=(B, H, W))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * B * H * W * Cout * 9 * Cin
    print(f"conv3x3 {H}x{W} {Cin}->{Cout} (M{B*H*W} N{Cout} K{9*Cin}): {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s")
