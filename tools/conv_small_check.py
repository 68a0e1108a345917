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
for (B, H, W, Cin, Cout) in [(2, 8, 8, 1280, 1280), (2, 8, 8, 2560, 1280), (2, 16, 16, 1280, 1280), (2, 16, 16, 2560, 1280), (2, 32, 32, 640, 640), (2, 32, 32, 1920, 640)]:
    x = torch.randn(B * H * W, Cin, device=DEV).bfloat16()
    wc = (torch.randn(Cout, 9 * Cin, device=DEV) / (9 * Cin) ** 0.5).bfloat16()
    o = torch.randn(B * H * W, Cout, device=DEV)
    bias = torch.randn(Cout, device=DEV)
    row = f"conv {B}x{H}x{W} {Cin}->{Cout}:"
    for bn, sk in ((0, 0), (160, 8), (160, 4), (80, 8), (80, 4), (80, 2), (160, 2), (64, 4), (128, 4)):
        try:
            args = ops.gemm(x, wc, o, conv=(B, H, W), bias=bias, residual=o, block_n=bn, split_k=sk, launch=False)
            t = timeit(lambda: ops.gemm_run(args))
            row += f" [bn{bn} sk{sk}] {t:5.1f}"
        except Exception as e:
            row += f" [bn{bn} sk{sk}] err"
    print(row, flush=True)
