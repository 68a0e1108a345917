"""Run each kernel repeatedly on identical inputs and report the max difference between runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
from b200sd.packing import pack_conv3x3, pack_geglu
DEV = "cuda:0"
torch.manual_seed(0)

def rep(name, fn, n=6):
    outs = [fn().clone() for _ in range(n)]
    torch.cuda.synchronize()
    d = max(float((o.float() - outs[0].float()).abs().max()) for o in outs[1:])
    print(f"{name:44s} max run-to-run diff {d:.3e}   (|out| max {float(outs[0].float().abs().max()):.3f})")

def gemm_case(M, N, K, split, residual_inplace=False):
    a = torch.randn(M, K, device=DEV).bfloat16(); w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV); out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    res = torch.randn(M, N, device=DEV).bfloat16()
    def fn():
        if residual_inplace:
            out.copy_(res)
            return ops.gemm(a, w, out, bias=bias, residual=out, split_k=split)
        return ops.gemm(a, w, out, bias=bias, residual=res, split_k=split)
    rep(f"gemm M{M} N{N} K{K} split{split} inplace{int(residual_inplace)}", fn)

for args in [(8192, 320, 320, 1), (8192, 320, 320, 1, True), (512, 1280, 1280, 0), (512, 1280, 1280, 0, True), (128, 1280, 1280, 0),
             (2048, 640, 640, 1), (2048, 640, 2560, 0), (512, 1280, 5120, 0, True), (8192, 960, 320, 1), (154, 640, 768, 1)]:
    gemm_case(*args)

def conv_case(B, H, W, Cin, Cout, C1=0):
    x = torch.randn(B * H * W, Cin, device=DEV).bfloat16()
    x1 = torch.randn(B * H * W, C1, device=DEV).bfloat16() if C1 else None
    w = pack_conv3x3((torch.randn(Cout, Cin + C1, 3, 3, device=DEV) / (9 * (Cin + C1)) ** 0.5))
    bias = torch.randn(Cout, device=DEV); out = torch.empty(B * H * W, Cout, device=DEV, dtype=torch.bfloat16)
    rep(f"conv B{B} {H}x{W} {Cin}+{C1}->{Cout}", lambda: ops.gemm(x, w, out, a1=x1, bias=bias, conv=(B, H, W)))
for args in [(2, 64, 64, 320, 320), (2, 32, 32, 640, 640), (2, 16, 16, 1280, 1280), (2, 8, 8, 1280, 1280), (2, 8, 8, 1280, 1280, 1280), (2, 16, 16, 1280, 1280, 640)]:
    conv_case(*args)

x = torch.randn(2 * 1024, 640, device=DEV).bfloat16(); w = torch.randn(5120, 640, device=DEV) / 25; b = torch.randn(5120, device=DEV)
tile = ops.geglu_tile(5120); wp, bp = pack_geglu(w, b, tile); out = torch.empty(2048, 2560, device=DEV, dtype=torch.bfloat16)
rep("gemm geglu", lambda: ops.gemm(x, wp, out, bias=bp, epilogue=ops.EPI_GEGLU, block_n=tile))

for (B, hw, C0, C1) in [(2, 4096, 320, 0), (2, 1024, 640, 320), (2, 64, 1280, 1280), (2, 256, 1280, 640)]:
    x0 = torch.randn(B * hw, C0, device=DEV).bfloat16(); x1 = torch.randn(B * hw, C1, device=DEV).bfloat16() if C1 else None
    g = torch.randn(C0 + C1, device=DEV); bb = torch.randn(C0 + C1, device=DEV); o = torch.empty(B * hw, C0 + C1, device=DEV, dtype=torch.bfloat16)
    rep(f"groupnorm B{B} hw{hw} C{C0}+{C1}", lambda: ops.groupnorm_silu(x0, x1, g, bb, o, B, hw, 32, 1e-5, True))

for (S, d) in [(4096, 40), (1024, 80), (256, 160)]:
    C = 8 * d
    qkv = torch.randn(2 * S, 3 * C, device=DEV).bfloat16(); o = torch.empty(2 * S, C, device=DEV, dtype=torch.bfloat16)
    rep(f"attention S{S} d{d}", lambda: ops.attention(qkv, qkv, qkv, o, 2, 8, S, S, d, d ** -0.5, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, k_off=C, v_off=2 * C))
x = torch.randn(8192, 320, device=DEV).bfloat16(); g = torch.randn(320, device=DEV); o = torch.empty_like(x)
rep("layernorm", lambda: ops.layernorm(x, g, g, o))
