"""GPU parity of the backward GEMMs (dgrad / wgrad, plain and 3x3 conv) against fp32 torch autograd
math on the same bf16-rounded operands.  These kernels read their operands MN-major (no transposed
copies), so a wrong UMMA descriptor shows up as O(1) errors, not rounding noise.
Tolerances: fp32 outputs 1e-4 of the tensor's max (fp32 accumulation order only); bf16 outputs 2^-7."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _close(got, want, rel):
    got, want = got.float(), want.float()
    scale = float(want.abs().max()) + 1e-6
    err = float((got - want).abs().max())
    assert err <= rel * scale, f"max abs err {err:.4g} vs scale {scale:.4g} (rel {err / scale:.3g})"


def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.mark.parametrize("M,Cout,Cin", [(128, 64, 64), (256, 320, 320), (8192, 960, 320), (616, 640, 768),
                                        (2048, 1280, 2560), (100, 2560, 320), (4096, 320, 1280)])
def test_dgrad_plain(M, Cout, Cin):
    from b200sd import ops
    _setup()
    torch.manual_seed(M + Cout + Cin)
    dy = torch.randn(M, Cout, device=DEV).bfloat16()
    w = (torch.randn(Cout, Cin, device=DEV) / Cout ** 0.5).bfloat16()
    want = dy.float() @ w.float()
    out = torch.empty(M, Cin, device=DEV, dtype=torch.float32)
    ops.gemm_dgrad(dy, w, out)
    _close(out, want, 1e-4)
    # bf16 output + accumulate into an fp32 running gradient
    outb = torch.empty(M, Cin, device=DEV, dtype=torch.bfloat16)
    ops.gemm_dgrad(dy, w, outb)
    _close(outb, want, 1.0 / 128)
    acc = torch.randn(M, Cin, device=DEV)
    want_acc = acc + want
    ops.gemm_dgrad(dy, w, acc, residual=acc)
    _close(acc, want_acc, 1e-4)


@pytest.mark.parametrize("bn", [64, 128, 192, 256])
def test_dgrad_tile_widths(bn):
    from b200sd import ops
    _setup()
    dy = torch.randn(384, 640, device=DEV).bfloat16()
    w = (torch.randn(640, 960, device=DEV) / 25).bfloat16()
    out = torch.empty(384, 960, device=DEV, dtype=torch.float32)
    ops.gemm_dgrad(dy, w, out, block_n=bn)
    _close(out, dy.float() @ w.float(), 1e-4)


@pytest.mark.parametrize("B,H,W,Cout,Cin", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 320), (3, 16, 16, 1280, 640),
                                            (2, 8, 8, 1280, 2560), (4, 4, 4, 64, 128), (1, 96, 64, 320, 640)])
def test_dgrad_conv3x3(B, H, W, Cout, Cin):
    from b200sd import ops
    _setup()
    torch.manual_seed(B + H + Cout + Cin)
    dy = torch.randn(B, H, W, Cout, device=DEV).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device=DEV) / (9 * Cout) ** 0.5).bfloat16()
    wp = w.permute(0, 2, 3, 1).contiguous().reshape(Cout, 9 * Cin)         # forward packing [Cout][ky][kx][Cin]
    want = torch.nn.grad.conv2d_input((B, Cin, H, W), w.float(), dy.float().permute(0, 3, 1, 2), padding=1)
    want = want.permute(0, 2, 3, 1).reshape(B * H * W, Cin)
    out = torch.empty(B * H * W, Cin, device=DEV, dtype=torch.float32)
    ops.gemm_dgrad(dy.reshape(B * H * W, Cout), wp, out, conv=(B, H, W))
    _close(out, want, 1e-4)


@pytest.mark.parametrize("rows,Cout,Cin", [(64, 64, 64), (128, 128, 64), (8192, 320, 320), (616, 1280, 768),
                                           (2048, 2560, 320), (1000, 320, 1280), (32768, 320, 320)])
def test_wgrad_plain(rows, Cout, Cin):
    from b200sd import ops
    _setup()
    torch.manual_seed(rows + Cout + Cin)
    dy = torch.randn(rows, Cout, device=DEV).bfloat16()
    x = torch.randn(rows, Cin, device=DEV).bfloat16()
    want = dy.float().t() @ x.float()
    dw = torch.zeros(Cout, Cin, device=DEV)
    ops.gemm_wgrad(dy, x, dw)
    _close(dw, want, 2e-4)
    ops.gemm_wgrad(dy, x, dw)            # accumulates
    _close(dw, 2 * want, 2e-4)


@pytest.mark.parametrize("split,bn", [(1, 64), (3, 128), (7, 192), (16, 256)])
def test_wgrad_split_and_tiles(split, bn):
    from b200sd import ops
    _setup()
    dy = torch.randn(4096, 384, device=DEV).bfloat16()
    x = torch.randn(4096, 960, device=DEV).bfloat16()
    dw = torch.zeros(384, 960, device=DEV)
    ops.gemm_wgrad(dy, x, dw, split_k=split, block_n=bn)
    _close(dw, dy.float().t() @ x.float(), 2e-4)


@pytest.mark.parametrize("B,H,W,Cout,Cin", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 320), (3, 16, 16, 1280, 640),
                                            (2, 8, 8, 1280, 2560), (5, 4, 4, 64, 128), (1, 96, 64, 320, 640)])
def test_wgrad_conv3x3(B, H, W, Cout, Cin):
    from b200sd import ops
    _setup()
    torch.manual_seed(B + H + Cout + Cin)
    dy = torch.randn(B, H, W, Cout, device=DEV).bfloat16()
    x = torch.randn(B, H, W, Cin, device=DEV).bfloat16()
    want = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, 3, 3), dy.float().permute(0, 3, 1, 2),
                                       padding=1)
    want = want.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin)                 # packed [Cout][ky][kx][Cin]
    dw = torch.zeros(Cout, 9 * Cin, device=DEV)
    ops.gemm_wgrad(dy.reshape(B * H * W, Cout), x.reshape(B * H * W, Cin), dw, conv=(B, H, W))
    _close(dw, want, 2e-4)


def test_forward_gemm_unchanged_by_bwd_modes():
    """The forward K-major path shares the kernel: a plain + conv sanity check guards the refactor."""
    from b200sd import ops
    _setup()
    a = torch.randn(512, 640, device=DEV).bfloat16()
    w = (torch.randn(320, 640, device=DEV) / 25).bfloat16()
    out = torch.empty(512, 320, device=DEV, dtype=torch.float32)
    ops.gemm(a, w, out)
    _close(out, a.float() @ w.float().t(), 1e-4)
    B, H, W, Cin, Cout = 2, 16, 16, 128, 64
    x = torch.randn(B, H, W, Cin, device=DEV).bfloat16()
    wc = (torch.randn(Cout, Cin, 3, 3, device=DEV) / 30).bfloat16()
    o = torch.empty(B * H * W, Cout, device=DEV, dtype=torch.float32)
    ops.gemm(x.reshape(-1, Cin), wc.permute(0, 2, 3, 1).contiguous().reshape(Cout, -1), o, conv=(B, H, W))
    want = F.conv2d(x.float().permute(0, 3, 1, 2), wc.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    _close(o, want, 1e-4)
