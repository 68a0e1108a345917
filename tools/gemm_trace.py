"""Timeline of the GEMM kernel's phases (globaltimer stamps written by every CTA): where do the
microseconds of a small GEMM go?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
from b200sd._lib import lib
from b200sd.packing import pack_conv3x3
DEV = "cuda:0"
L = lib()
trace = torch.zeros(4096 * 16, dtype=torch.int64, device=DEV)
names = ["start", "prologue done", "first tile landed", "all MMA issued", "accum ready (epi)", "phaseA done | chunks issued", "phaseB done | stores drained", "exit"]
# persistent kernel: slot 5 = warp 2 issued its last TMA store, slot 6 = its stores have completed

def run(label, fn, nctas):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    trace.zero_()
    L.b200sd_debug_gemm_trace(trace.data_ptr())
    fn(); torch.cuda.synchronize()
    L.b200sd_debug_gemm_trace(None)
    # 16 slots per CTA: 8 phase times (+ for the persistent kernel 6 epilogue phase tick sums of warp 2)
    t16 = trace[: nctas * 16].view(nctas, 16).cpu().double()
    t = t16[:, :8]
    persist = bool((t16[:, 8:14] != 0).any())
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    print(f"== {label}: {nctas} CTAs; kernel span {float(rel[:, 7].max()):.2f} us")
    print("   phase                  median   min     max  (us since first CTA start)")
    for i, n in enumerate(names):
        c = rel[:, i]
        print(f"   {n:20s} {float(c.median()):7.2f} {float(c.min()):7.2f} {float(c.max()):7.2f}")
    if persist:
        lab = ["bias staging", "wait accumulator", "wait smem chunk", "TMEM->regs->math->smem", "fence + TMA store issue", "loop overhead"]
        ticks = t16[:, 8:14].median(dim=0).values
        print("   epilogue of warp 2 (median over CTAs, us at 1.965 GHz): " + ", ".join(f"{l} {float(v) / 1965.0:.2f}" for l, v in zip(lab, ticks)))

def gemm_case(M, N, K, res=True, f32=True):
    a = torch.randn(M, K, device=DEV).bfloat16(); w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV)
    out = torch.randn(M, N, device=DEV) if f32 else torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    r = out if res else None
    args = ops.gemm(a, w, out, bias=bias, residual=r, launch=False)
    return (lambda: ops.gemm_run(args)), (a, w, bias, out)

for (M, N, K, res, f32) in [(8192, 320, 320, True, True), (8192, 320, 320, False, False), (8192, 960, 320, False, False), (512, 1280, 1280, True, True), (2048, 640, 640, True, True), (8192, 320, 1280, True, True)]:
    fn, keep = gemm_case(M, N, K, res, f32)
    # ask the library how many CTAs: replicate heuristics crudely by reading the trace afterwards
    trace.zero_(); L.b200sd_debug_gemm_trace(trace.data_ptr()); fn(); torch.cuda.synchronize(); L.b200sd_debug_gemm_trace(None)
    n = int((trace.view(-1, 16)[:, 0] != 0).sum())
    run(f"gemm M{M} N{N} K{K} (+bias, residual={res}, fp32 out={f32})", fn, n)

# GEGLU feed-forward (M8192 N2560 K320): multi-tile persistent CTAs
for (M, N, K) in [(8192, 2560, 320), (2048, 5120, 640)]:
    a = torch.randn(M, K, device=DEV).bfloat16(); w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV); out = torch.empty(M, N // 2, device=DEV, dtype=torch.bfloat16)
    argsg = ops.gemm(a, w, out, bias=bias, epilogue=ops.EPI_GEGLU, block_n=ops.geglu_tile(N), launch=False)
    fng = lambda: ops.gemm_run(argsg)
    trace.zero_(); L.b200sd_debug_gemm_trace(trace.data_ptr()); fng(); torch.cuda.synchronize(); L.b200sd_debug_gemm_trace(None)
    run(f"GEGLU gemm M{M} N{N} K{K} tile {ops.geglu_tile(N)}", fng, int((trace.view(-1, 16)[:, 0] != 0).sum()))

B, H, W, Cin, Cout = 2, 8, 8, 1280, 1280
x = torch.randn(B * H * W, Cin, device=DEV).bfloat16()
w = pack_conv3x3(torch.randn(Cout, Cin, 3, 3, device=DEV) / (9 * Cin) ** 0.5); bias = torch.randn(Cout, device=DEV)
out = torch.randn(B * H * W, Cout, device=DEV)
args = ops.gemm(x, w, out, bias=bias, residual=out, conv=(B, H, W), launch=False)
fn = lambda: ops.gemm_run(args)
trace.zero_(); L.b200sd_debug_gemm_trace(trace.data_ptr()); fn(); torch.cuda.synchronize(); L.b200sd_debug_gemm_trace(None)
run("conv3x3 8x8 1280->1280 (M128 N1280 K11520)", fn, int((trace.view(-1, 16)[:, 0] != 0).sum()))
B, H, W, Cin, Cout = 2, 64, 64, 320, 320
x = torch.randn(B * H * W, Cin, device=DEV).bfloat16()
w = pack_conv3x3(torch.randn(Cout, Cin, 3, 3, device=DEV) / (9 * Cin) ** 0.5); bias = torch.randn(Cout, device=DEV)
out = torch.randn(B * H * W, Cout, device=DEV)
args2 = ops.gemm(x, w, out, bias=bias, residual=out, conv=(B, H, W), launch=False)
fn2 = lambda: ops.gemm_run(args2)
trace.zero_(); L.b200sd_debug_gemm_trace(trace.data_ptr()); fn2(); torch.cuda.synchronize(); L.b200sd_debug_gemm_trace(None)
run("conv3x3 64x64 320->320 (M8192 N320 K2880)", fn2, int((trace.view(-1, 16)[:, 0] != 0).sum()))
