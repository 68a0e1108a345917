import sys, os, ctypes; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd._lib import lib
from b200sd import ops
L = lib()
L.b200sd_debug_gemm_trace.argtypes = [ctypes.c_void_p]
trace = torch.zeros(4096 * 8, dtype=torch.int64, device='cuda')
names = ["start", "prologue done", "first tile landed", "all MMA issued", "accum ready (epi)", "phaseA done", "phaseB done", "exit"]
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
for (M, N, K, f32, res) in [(32768, 1280, 320, False, False), (32768, 320, 320, True, True), (8192, 320, 320, True, True), (8192, 2560, 320, False, False),
                            (32768, 320, 1280, True, True), (8192, 1280, 1280, False, False)]:
    a = torch.randn(M, K, device='cuda').bfloat16(); w = torch.randn(N, K, device='cuda').bfloat16()
    out = torch.empty(M, N, device='cuda', dtype=torch.float32 if f32 else torch.bfloat16)
    args = ops.gemm(a, w, out, launch=False, residual=out if res else None)
    t_us = timeit(lambda: ops.gemm_run(args))
    torch.cuda.synchronize(); trace.zero_(); L.b200sd_debug_gemm_trace(trace.data_ptr()); ops.gemm_run(args); torch.cuda.synchronize(); L.b200sd_debug_gemm_trace(None)
    t = trace.view(-1, 8)
    n = int((t[:, 0] != 0).sum()); t = t[:n].double().cpu()
    rel = (t - t[:, :1]) / 1e3
    print(f"M{M} N{N} K{K} f32out={f32} residual={res}: {t_us:.1f} us ({2*M*N*K/t_us/1e6:.0f} TF/s), {n} CTAs; per-CTA median us since its own start:")
    print("   " + "  ".join(f"{nm}={float(rel[:, i].median()):.2f}" for i, nm in enumerate(names) if i not in (6,)))
