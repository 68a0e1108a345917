import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
DEV = "cuda:0"
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=s):
        for _ in range(reps): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e3
for (B, H, W, Cin, Cout) in [(2, 8, 8, 1280, 1280), (2, 8, 8, 2560, 1280), (2, 16, 16, 1280, 1280), (2, 32, 32, 640, 640), (2, 32, 32, 1280, 640), (2, 64, 64, 320, 320)]:
    x = torch.randn(B * H * W, Cin, device=DEV).bfloat16()
    wc = (torch.randn(Cout, 9 * Cin, device=DEV) / (9 * Cin) ** 0.5).bfloat16()
    o = torch.randn(B * H * W, Cout, device=DEV)
    bias = torch.randn(Cout, device=DEV)
    args = ops.gemm(x, wc, o, conv=(B, H, W), bias=bias, residual=o, launch=False)
    t = timeit(lambda: ops.gemm_run(args))
    print(f"conv {B}x{H}x{W} {Cin}->{Cout}: {t:7.1f} us {2 * B * H * W * Cout * 9 * Cin / t / 1e6:6.0f} TF/s", flush=True)
