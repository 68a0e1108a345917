"""GPU parity of the tcgen05 GEMM / implicit-GEMM conv kernel against fp32 torch math on the same
bf16-rounded operands.  Tolerance: bf16 output rounding (2^-8 relative) + fp32 accumulation order."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _close(got, want, rel=1.0 / 128):
    got, want = got.float(), want.float()
    scale = float(want.abs().max()) + 1e-6
    err = float((got - want).abs().max())
    assert err <= rel * scale, f"max abs err {err:.4g} vs scale {scale:.4g} (rel {err / scale:.3g})"


def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.mark.parametrize("M,N,K", [(128, 160, 64), (128, 16, 128), (256, 320, 320), (8192, 320, 320),
                                   (154, 640, 768), (2048, 1920, 640), (100, 256, 1280), (8192, 640, 2560)])
def test_gemm_plain(M, N, K):
    from b200sd import ops
    _setup()
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV).bfloat16()
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=bias, residual=res)
    want = a.float() @ w.float().t() + bias + res.float()
    _close(out, want)


def test_gemm_f32_out_no_bias():
    from b200sd import ops
    _setup()
    a = torch.randn(512, 640, device=DEV).bfloat16()
    w = (torch.randn(320, 640, device=DEV) / 25).bfloat16()
    out = torch.empty(512, 320, device=DEV, dtype=torch.float32)
    ops.gemm(a, w, out)
    _close(out, a.float() @ w.float().t(), rel=1e-4)


@pytest.mark.parametrize("split", [2, 4, 8])
def test_gemm_split_k(split):
    from b200sd import ops
    _setup()
    M, N, K = 128, 1280, 5120
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    want = a.float() @ w.float().t() + bias
    for _ in range(2):  # twice: the self-cleaning tile counters must be reusable
        out.zero_()
        ops.gemm(a, w, out, bias=bias, split_k=split)
        _close(out, want)


def test_gemm_dual_source_concat():
    from b200sd import ops
    _setup()
    M, C0, C1, N = 2048, 640, 320, 640
    a0 = torch.randn(M, C0, device=DEV).bfloat16()
    a1 = torch.randn(M, C1, device=DEV).bfloat16()
    w = (torch.randn(N, C0 + C1, device=DEV) / 30).bfloat16()
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a0, w, out, a1=a1)
    _close(out, torch.cat([a0, a1], 1).float() @ w.float().t())


@pytest.mark.parametrize("C", [320, 640])
def test_gemm_geglu(C):
    from b200sd import ops
    from b200sd.packing import pack_geglu
    _setup()
    M = 1024
    x = torch.randn(M, C, device=DEV).bfloat16()
    w = (torch.randn(8 * C, C, device=DEV) / C ** 0.5)
    b = torch.randn(8 * C, device=DEV)
    tile = ops.geglu_tile(8 * C)
    wp, bp = pack_geglu(w, b, tile)
    out = torch.empty(M, 4 * C, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x, wp, out, bias=bp, epilogue=ops.EPI_GEGLU, block_n=tile)
    h = x.float() @ w.bfloat16().float().t() + b
    val, gate = h.chunk(2, dim=-1)
    _close(out, val * F.gelu(gate))


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 640), (2, 16, 16, 1280, 1280),
                                            (2, 8, 8, 1280, 1280), (1, 8, 8, 1280, 1280), (3, 8, 8, 640, 320),
                                            (1, 96, 64, 320, 320), (2, 12, 8, 1280, 1280), (1, 48, 32, 640, 640)])
def test_conv3x3(B, H, W, Cin, Cout):
    from b200sd import ops
    from b200sd.packing import pack_conv3x3
    _setup()
    torch.manual_seed(H * W + Cin)
    x = torch.randn(B, Cin, H, W, device=DEV).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device=DEV) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device=DEV)
    temb = torch.randn(B, Cout, device=DEV)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().reshape(B * H * W, Cin)
    out = torch.empty(B * H * W, Cout, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x_nhwc, pack_conv3x3(w), out, bias=bias, rowbias=temb, conv=(B, H, W), rows_per_image=H * W)
    want = F.conv2d(x.float(), w.float(), bias, padding=1) + temb[:, :, None, None]
    want = want.permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    _close(out, want)


def test_conv3x3_dual_source_residual():
    from b200sd import ops
    from b200sd.packing import pack_conv3x3
    _setup()
    B, H, W, C0, C1, Cout = 2, 16, 16, 1280, 640, 1280
    x0 = torch.randn(B, C0, H, W, device=DEV).bfloat16()
    x1 = torch.randn(B, C1, H, W, device=DEV).bfloat16()
    w = (torch.randn(Cout, C0 + C1, 3, 3, device=DEV) / (9 * (C0 + C1)) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device=DEV)
    res = torch.randn(B * H * W, Cout, device=DEV).bfloat16()
    n0 = x0.permute(0, 2, 3, 1).contiguous().reshape(-1, C0)
    n1 = x1.permute(0, 2, 3, 1).contiguous().reshape(-1, C1)
    out = torch.empty(B * H * W, Cout, device=DEV, dtype=torch.bfloat16)
    ops.gemm(n0, pack_conv3x3(w), out, a1=n1, bias=bias, residual=res, conv=(B, H, W))
    want = F.conv2d(torch.cat([x0, x1], 1).float(), w.float(), bias, padding=1)
    want = want.permute(0, 2, 3, 1).reshape(B * H * W, Cout) + res.float()
    _close(out, want)


def test_gemm_rejects_bad_shapes():
    from b200sd import ops
    from b200sd._lib import B200SDError
    a = torch.randn(128, 100, device=DEV).bfloat16()  # K not a multiple of 64
    w = torch.randn(64, 100, device=DEV).bfloat16()
    out = torch.empty(128, 64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(B200SDError):
        ops.gemm(a, w, out)


@pytest.mark.parametrize("M,N,K,split", [(8192, 320, 320, 1), (512, 1280, 5120, 0), (128, 1280, 2560, 4)])
def test_gemm_fp32_residual_and_output(M, N, K, split):
    """fp32 residual stream: out(fp32) = A W^T + bias + residual(fp32), incl. in-place (out is residual)."""
    from b200sd import ops
    _setup()
    torch.manual_seed(M + K)
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    want = a.float() @ w.float().t() + bias + res
    out = res.clone()
    ops.gemm(a, w, out, bias=bias, residual=out, split_k=split)
    _close(out, want, rel=2e-5)
    outb = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, outb, bias=bias, residual=res, split_k=split)
    _close(outb, want)
