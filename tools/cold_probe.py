"""Probe with COLD weights: every launch of the graph reads a different copy of the weight matrix (copies total > 2x the 126 MB
L2), the activations stay warm -- the condition of a GEMM inside the real step.  us per launch: old kernel vs persistent."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
from b200sd._lib import lib
DEV = "cuda:0"
L = lib()

def timeit(fns):
    for f in fns[:3]: f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=s):
        for f in fns: f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / len(fns))
    return best * 1e3

def case(kind, M, N, K, conv=None, res=True, f32=True, geglu=False, variants=((0, 0), (1, 0))):
    C = K // 9 if conv else K
    ncopy = max(8, int(300e6 / (N * K * 2)) + 1)
    ncopy = min(ncopy, 400)
    a = torch.randn(M, C, device=DEV).bfloat16()
    ws = [(torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16() for _ in range(ncopy)]
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
    row = f"{kind} M{M} N{N} K{K} ({ncopy} weight copies):"
    for persist, bn in variants:
        L.b200sd_debug_set(0, persist)
        b = bn if not geglu else (bn or (256 if persist else 160))
        try:
            argl = [ops.gemm(a, w, out, block_n=b, launch=False, **kw) for w in ws]
            fns = [(lambda x=x: ops.gemm_run(x)) for x in argl]
            row += f" [{'P' if persist else 'o'} bn{b}] {timeit(fns):5.1f}"
        except Exception as e:
            row += f" [{'P' if persist else 'o'} bn{b}] err {str(e)[:40]}"
    L.b200sd_debug_set(0, -1)
    print(row, flush=True)

if __name__ == "__main__": case("conv", 128, 1280, 11520, conv=(2, 8, 8))
if __name__ == "__main__": case("conv", 128, 1280, 23040, conv=(2, 8, 8))
if __name__ == "__main__": case("conv", 512, 1280, 11520, conv=(2, 16, 16))
if __name__ == "__main__": case("conv", 512, 1280, 23040, conv=(2, 16, 16))
if __name__ == "__main__": case("gemm", 512, 1280, 1280)
if __name__ == "__main__": case("gemm", 512, 1280, 5120)
if __name__ == "__main__": case("gemm", 2048, 640, 2560)
if __name__ == "__main__": case("gemm", 128, 1280, 1280)
if __name__ == "__main__": case("conv", 2048, 640, 5760, conv=(2, 32, 32))
if __name__ == "__main__" and os.environ.get("SMALL_ONLY"): sys.exit(0)
if __name__ == "__main__": case("gemm", 8192, 320, 320)
if __name__ == "__main__": case("gemm", 2048, 640, 640)
if __name__ == "__main__": case("gemm", 8192, 320, 1280)
if __name__ == "__main__": case("gemm", 8192, 960, 320, res=False, f32=False)
if __name__ == "__main__": case("gemm", 2048, 1920, 640, res=False, f32=False)
if __name__ == "__main__": case("conv", 8192, 320, 2880, conv=(2, 64, 64))
if __name__ == "__main__": case("conv", 2048, 640, 5760, conv=(2, 32, 32))
if __name__ == "__main__": case("geglu", 8192, 2560, 320, geglu=True, variants=((0, 160), (1, 256), (1, 128)))
if __name__ == "__main__": case("geglu", 2048, 5120, 640, geglu=True, variants=((0, 160), (1, 256), (1, 128)))
if __name__ == "__main__": case("geglu", 512, 10240, 1280, geglu=True, variants=((0, 160), (1, 256), (1, 128)))
