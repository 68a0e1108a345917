"""CTA-pair (cta_group::2) GEMM mode: correctness vs torch and A/B timing against the single-CTA tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from b200sd import ops
DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rel(got, want):
    return float((got.float() - want.float()).abs().max() / (want.float().abs().max() + 1e-9))


def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=s):
        for _ in range(reps): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e3


print("== correctness (pair=1) ==")
for (M, N, K, bn) in [(256, 160, 64, 0), (256, 320, 320, 0), (384, 320, 640, 0), (8192, 320, 320, 0), (8192, 1280, 320, 256), (300, 640, 768, 0),
                      (2048, 1920, 640, 192), (32768, 320, 1280, 0)]:
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV).bfloat16(); w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV); res = torch.randn(M, N, device=DEV)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(a, w, out, bias=bias, residual=res, pair=1, split_k=1, block_n=bn)
    torch.cuda.synchronize()
    print(f"gemm M{M} N{N} K{K} bn{bn}: rel {rel(out, a.float() @ w.float().t() + bias + res):.3g}", flush=True)
for (B, H, W, Cin, Cout) in [(2, 64, 64, 320, 320), (3, 16, 16, 640, 1280), (2, 8, 8, 1280, 1280), (8, 32, 32, 640, 640)]:
    x = torch.randn(B, H, W, Cin, device=DEV).bfloat16()
    wc = (torch.randn(Cout, Cin, 3, 3, device=DEV) / (9 * Cin) ** 0.5).bfloat16()
    o = torch.empty(B * H * W, Cout, device=DEV)
    ops.gemm(x.reshape(-1, Cin), wc.permute(0, 2, 3, 1).contiguous().reshape(Cout, -1), o, conv=(B, H, W), pair=1, split_k=1)
    torch.cuda.synchronize()
    want = F.conv2d(x.float().permute(0, 3, 1, 2), wc.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    print(f"conv {B}x{H}x{W} {Cin}->{Cout}: rel {rel(o, want):.3g}", flush=True)
for (M, Cout, Cin) in [(256, 320, 320), (8192, 960, 320), (2048, 1280, 2560)]:
    dy = torch.randn(M, Cout, device=DEV).bfloat16(); w = (torch.randn(Cout, Cin, device=DEV) / Cout ** 0.5).bfloat16()
    out = torch.empty(M, Cin, device=DEV)
    ops.gemm_dgrad(dy, w, out, pair=1)
    torch.cuda.synchronize()
    print(f"dgrad M{M} Cout{Cout} Cin{Cin}: rel {rel(out, dy.float() @ w.float()):.3g}", flush=True)
B, H, W, Cout, Cin = 2, 32, 32, 640, 320
dy = torch.randn(B, H, W, Cout, device=DEV).bfloat16(); w = (torch.randn(Cout, Cin, 3, 3, device=DEV) / (9 * Cout) ** 0.5).bfloat16()
out = torch.empty(B * H * W, Cin, device=DEV)
ops.gemm_dgrad(dy.reshape(-1, Cout), w.permute(0, 2, 3, 1).contiguous().reshape(Cout, -1), out, conv=(B, H, W), pair=1)
want = torch.nn.grad.conv2d_input((B, Cin, H, W), w.float(), dy.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1).reshape(-1, Cin)
print(f"dgrad conv: rel {rel(out, want):.3g}", flush=True)

print("== timing: single-CTA tiles vs CTA pairs (us, TF/s) ==")
for (M, N, K) in [(8192, 320, 320), (8192, 320, 1280), (8192, 2560, 320), (32768, 320, 320), (32768, 1280, 320), (32768, 320, 1280), (8192, 1280, 1280),
                  (16384, 1280, 1280), (2048, 640, 640), (512, 1280, 1280)]:
    a = torch.randn(M, K, device=DEV).bfloat16(); w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    row = f"gemm M{M} N{N} K{K}:"
    for pair, bn in ((-1, 0), (1, 0), (1, 256 if N % 256 == 0 else (160 if N % 160 == 0 else 0)), (1, 128 if N % 128 == 0 else 0)):
        args = ops.gemm(a, w, out, launch=False, pair=pair, block_n=bn, split_k=1 if pair > 0 else 0)
        t = timeit(lambda: ops.gemm_run(args))
        row += f"  [pair {pair} bn {bn}] {t:7.1f} us {2 * M * N * K / t / 1e6:6.0f} TF/s"
    print(row, flush=True)
for (B, H, W, Cin, Cout) in [(2, 64, 64, 320, 320), (8, 64, 64, 320, 320), (8, 32, 32, 640, 640), (8, 16, 16, 1280, 1280), (2, 32, 32, 640, 640), (8, 64, 64, 640, 320)]:
    x = torch.randn(B * H * W, Cin, device=DEV).bfloat16()
    wc = (torch.randn(Cout, 9 * Cin, device=DEV) / (9 * Cin) ** 0.5).bfloat16()
    o = torch.empty(B * H * W, Cout, device=DEV)
    row = f"conv {B}x{H}x{W} {Cin}->{Cout}:"
    for pair, bn in ((-1, 0), (1, 0), (1, 256 if Cout % 256 == 0 else 160), (1, 128 if Cout % 128 == 0 else 0)):
        args = ops.gemm(x, wc, o, conv=(B, H, W), launch=False, pair=pair, block_n=bn, split_k=1 if pair > 0 else 0)
        t = timeit(lambda: ops.gemm_run(args))
        row += f"  [pair {pair} bn {bn}] {t:7.1f} us {2 * B * H * W * Cout * 9 * Cin / t / 1e6:6.0f} TF/s"
    print(row, flush=True)
