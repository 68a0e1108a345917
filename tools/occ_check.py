import sys, os, ctypes; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd._lib import lib
from b200sd import ops
L = lib()
L.b200sd_debug_gemm_trace.argtypes = [ctypes.c_void_p]
trace = torch.zeros(4096 * 8, dtype=torch.int64, device='cuda')
names = ["start", "prologue done", "first tile landed", "all MMA issued", "accum ready (epi)", "phaseA done", "phaseB done", "exit"]
for (M, N, K, f32, res) in [(32768, 1280, 320, False, False), (32768, 320, 320, True, True), (8192, 2560, 320, False, False)]:
    a = torch.randn(M, K, device='cuda').bfloat16(); w = torch.randn(N, K, device='cuda').bfloat16()
    out = torch.empty(M, N, device='cuda', dtype=torch.float32 if f32 else torch.bfloat16)
    args = ops.gemm(a, w, out, launch=False, pair=-1, split_k=1, residual=out if res else None)
    for _ in range(3): ops.gemm_run(args)
    torch.cuda.synchronize(); trace.zero_(); L.b200sd_debug_gemm_trace(trace.data_ptr()); ops.gemm_run(args); torch.cuda.synchronize(); L.b200sd_debug_gemm_trace(None)
    t = trace.view(-1, 8)
    n = int((t[:, 0] != 0).sum()); t = t[:n].double().cpu()
    rel = (t - t[:, :1]) / 1e3      # per-CTA phases relative to that CTA's own start
    print(f"M{M} N{N} K{K} f32out={f32} residual={res}: {n} CTAs, span {float((t[:,7].max()-t[:,0].min())/1e3):.1f} us; per-CTA median us since its own start:")
    print("   " + "  ".join(f"{nm}={float(rel[:, i].median()):.2f}" for i, nm in enumerate(names)))
