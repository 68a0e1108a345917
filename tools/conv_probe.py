"""Tuning probe: a few hot shapes of the CFG-batch-2 step under {old kernel, persistent kernel} x tile width x ring depth.
Warm, 20 launches per CUDA graph, best of 3 (us per launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
from b200sd._lib import lib
DEV = "cuda:0"
L = lib()

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

def case(kind, M, N, K, conv=None, res=True, f32=True, geglu=False):
    C = K // 9 if conv else K
    a = torch.randn(M, C, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV)
    if geglu:
        out = torch.empty(M, N // 2, device=DEV, dtype=torch.bfloat16)
    else:
        out = torch.randn(M, N, device=DEV) if f32 else torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    kw = dict(bias=bias, conv=conv)
    if geglu:
        kw.update(epilogue=ops.EPI_GEGLU)
    elif res:
        kw.update(residual=out)
    row = f"{kind} M{M} N{N} K{K}:"
    for persist, bn, st in ((0, 0, 0), (1, 0, 0), (1, 0, 2), (1, 0, 3), (1, 64, 0), (1, 96, 0), (1, 128, 0), (1, 160, 0), (1, 256, 0), (0, 80, 0), (0, 160, 3)):
        if geglu and bn not in (0, 128, 256):
            continue
        L.b200sd_debug_set(0, persist); L.b200sd_debug_set(1, st)
        try:
            b = bn if not geglu else (bn or ops.geglu_tile(N))
            args = ops.gemm(a, w, out, block_n=b, launch=False, **kw)
            t = timeit(lambda: ops.gemm_run(args))
            row += f" [{'P' if persist else 'o'} bn{bn} st{st}] {t:5.1f}"
        except Exception as e:
            row += f" [{'P' if persist else 'o'} bn{bn} st{st}] err"
    L.b200sd_debug_set(0, -1); L.b200sd_debug_set(1, 0)
    print(row, flush=True)

case("conv", 8192, 320, 2880, conv=(2, 64, 64))
case("conv", 8192, 320, 5760, conv=(2, 64, 64))
case("conv", 2048, 640, 5760, conv=(2, 32, 32))
case("gemm", 8192, 320, 320)
case("gemm", 8192, 320, 1280)
case("gemm", 2048, 640, 640)
case("gemm", 8192, 960, 320, res=False, f32=False)
case("gemm", 2048, 1920, 640, res=False, f32=False)
case("geglu", 8192, 2560, 320, geglu=True)
case("geglu", 2048, 5120, 640, geglu=True)
case("geglu", 512, 10240, 1280, geglu=True)
