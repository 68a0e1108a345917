"""GPU parity of the non-GEMM UNet building-block kernels against fp32 torch math."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def _close(got, want, rel):
    got, want = got.float(), want.float()
    scale = float(want.abs().max()) + 1e-6
    err = float((got - want).abs().max())
    assert err <= rel * scale, f"max abs err {err:.4g} vs scale {scale:.4g} (rel {err / scale:.3g})"


def test_timestep_embedding_kat():
    from b200sd import ops
    from oracle.unet_ref import timestep_embedding
    t = torch.tensor([980.0, 1.0, 500.0, 0.0, 999.0], device=DEV)
    got = ops.timestep_embedding(t, 320).cpu()
    want = timestep_embedding(t.cpu().long(), 320)
    torch.testing.assert_close(got, want, rtol=0, atol=2e-4)  # fp32 sin/cos of ~1e3 rad arguments
    # SURVEY.md App. B.5 known answers
    torch.testing.assert_close(got[0, :3], torch.tensor([0.98439258, 0.01940880, 0.99800313]), rtol=0, atol=2e-4)
    torch.testing.assert_close(got[0, 160:163], torch.tensor([-0.17598660, 0.99981165, 0.06316458]), rtol=0, atol=2e-4)


@pytest.mark.parametrize("B,N,K", [(2, 1280, 320), (2, 19200, 1280), (11, 640, 1280), (1, 320, 1280)])
def test_small_linear(B, N, K):
    from b200sd import ops
    torch.manual_seed(0)
    x = torch.randn(B, K, device=DEV)
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=DEV)
    got = ops.small_linear(x, w, b, silu_in=True, silu_out=False)
    want = F.silu(x) @ w.float().t() + b
    _close(got, want, 1e-5)
    got = ops.small_linear(x, w, b, silu_in=False, silu_out=True)
    _close(got, F.silu(x @ w.float().t() + b), 1e-5)


@pytest.mark.parametrize("B,hw,C0,C1,silu,eps", [(2, 4096, 320, 0, True, 1e-5), (2, 1024, 640, 320, True, 1e-5),
                                                 (2, 64, 1280, 1280, True, 1e-5), (1, 256, 1280, 640, True, 1e-5),
                                                 (3, 1024, 640, 0, False, 1e-6), (2, 6144, 320, 320, True, 1e-5)])
def test_groupnorm(B, hw, C0, C1, silu, eps):
    from b200sd import ops
    torch.manual_seed(1)
    C = C0 + C1
    x0 = (torch.randn(B * hw, C0, device=DEV) * 2 + 0.5).bfloat16()
    x1 = (torch.randn(B * hw, C1, device=DEV) - 1.0).bfloat16() if C1 else None
    g = torch.randn(C, device=DEV)
    b = torch.randn(C, device=DEV)
    out = torch.empty(B * hw, C, device=DEV, dtype=torch.bfloat16)
    ops.groupnorm_silu(x0, x1, g, b, out, B, hw, 32, eps, silu)
    xc = x0 if x1 is None else torch.cat([x0, x1], 1)
    xn = xc.float().reshape(B, hw, C).permute(0, 2, 1)
    want = F.group_norm(xn, 32, g, b, eps)
    if silu:
        want = F.silu(want)
    want = want.permute(0, 2, 1).reshape(B * hw, C)
    _close(out, want, 1.0 / 128)


@pytest.mark.parametrize("rows,C", [(8192, 320), (2048, 640), (517, 1280), (3, 64)])
def test_layernorm(rows, C):
    from b200sd import ops
    torch.manual_seed(2)
    x = (torch.randn(rows, C, device=DEV) * 3 + 1).bfloat16()
    g = torch.randn(C, device=DEV)
    b = torch.randn(C, device=DEV)
    out = torch.empty_like(x)
    ops.layernorm(x, g, b, out)
    _close(out, F.layer_norm(x.float(), (C,), g, b, 1e-5), 1.0 / 128)


@pytest.mark.parametrize("B,heads,Sq,Skv,d", [(2, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80), (2, 8, 256, 256, 160),
                                              (2, 8, 64, 64, 160), (2, 8, 4096, 77, 40), (2, 8, 1024, 77, 80),
                                              (1, 8, 64, 77, 160), (1, 2, 200, 130, 32), (1, 8, 6144, 6144, 40),
                                              (3, 8, 96, 77, 160)])
def test_attention(B, heads, Sq, Skv, d):
    from b200sd import ops
    torch.manual_seed(3)
    C = heads * d
    q = torch.randn(B * Sq, C, device=DEV).bfloat16()
    kv = torch.randn(B * Skv, 2 * C, device=DEV).bfloat16()  # fused [K | V] buffer, like the UNet's
    out = torch.empty(B * Sq, C, device=DEV, dtype=torch.bfloat16)
    scale = 2.0 * d ** -0.5  # sharper than default so the softmax is not flat
    ops.attention(q, kv, kv, out, B, heads, Sq, Skv, d, scale, ldk=2 * C, ldv=2 * C, k_off=0, v_off=C)
    qf = q.float().reshape(B, Sq, heads, d).transpose(1, 2)
    kf = kv[:, :C].float().reshape(B, Skv, heads, d).transpose(1, 2)
    vf = kv[:, C:].float().reshape(B, Skv, heads, d).transpose(1, 2)
    p = (qf @ kf.transpose(-1, -2) * scale).softmax(-1)
    want = (p @ vf).transpose(1, 2).reshape(B * Sq, C)
    _close(out, want, 1.0 / 64)


@pytest.mark.parametrize("B,S,d,ramp", [(2, 4096, 40, 6.0), (2, 1024, 80, 6.0), (1, 1024, 40, 30.0), (1, 192, 40, 0.0)])
def test_attention_growing_scores(B, S, d, ramp):
    """Scores that keep growing along the keys: the tcgen05 kernel exponentiates every tile against a lagging reference
    max and must redo the tile / rescale its TMEM-resident O rows when a row outgrows it; rows of different warps do
    so at different tiles.  Also checks the log-sum-exp handed to the backward.  (S=192: a 64-key tail tile.)"""
    from b200sd import ops
    torch.manual_seed(9)
    H = 8
    C = H * d
    qkv = torch.randn(B * S, 3 * C, device=DEV)
    if ramp > 0:
        qkv[:, C:2 * C] *= (1.0 + ramp * torch.arange(S, device=DEV).repeat(B) / S)[:, None]
    qkv = qkv.bfloat16()
    out = torch.empty(B * S, C, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device=DEV)
    ops.attention_lse(qkv, qkv, qkv, out, lse, B, H, S, S, d, d ** -0.5, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, k_off=C,
                      v_off=2 * C)
    q, k, v = [t.reshape(B, S, H, d).permute(0, 2, 1, 3).float() for t in qkv.split(C, dim=1)]
    sc = (q @ k.transpose(-1, -2)) * d ** -0.5
    want = (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, C)
    _close(out, want, 1.0 / 64)
    want_lse = torch.logsumexp(sc, -1) * 1.4426950408889634
    assert (lse - want_lse).abs().max().item() < 2e-3


def test_conv_in_out():
    from b200sd import ops
    from b200sd.packing import pack_conv3x3_f32
    torch.manual_seed(4)
    B, H, W = 2, 64, 64
    x = torch.randn(B, 4, H, W, device=DEV)
    w = torch.randn(320, 4, 3, 3, device=DEV) / 6
    b = torch.randn(320, device=DEV)
    out = torch.empty(B * H * W, 320, device=DEV, dtype=torch.bfloat16)
    ops.conv_in(x, pack_conv3x3_f32(w), b, out)
    want = F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1).reshape(B * H * W, 320)
    _close(out, want, 1.0 / 128)

    y = torch.randn(B * H * W, 320, device=DEV).bfloat16()
    w2 = torch.randn(4, 320, 3, 3, device=DEV) / 54
    b2 = torch.randn(4, device=DEV)
    o2 = torch.empty(B, 4, H, W, device=DEV)
    ops.conv_out(y, pack_conv3x3_f32(w2), b2, o2)
    yn = y.float().reshape(B, H, W, 320).permute(0, 3, 1, 2)
    _close(o2, F.conv2d(yn, w2, b2, padding=1), 1e-4)


def test_upsample_and_im2col():
    from b200sd import ops
    torch.manual_seed(5)
    B, H, W, C = 2, 16, 16, 1280
    x = torch.randn(B, H, W, C, device=DEV).bfloat16()
    up = torch.empty(B, 2 * H, 2 * W, C, device=DEV, dtype=torch.bfloat16)
    ops.upsample2x(x.reshape(-1, C), up.reshape(-1, C), B, H, W)
    want = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), want)
    col = torch.empty(B * (H // 2) * (W // 2), 9 * C, device=DEV, dtype=torch.bfloat16)
    ops.im2col_s2(x.reshape(-1, C), col, B, H, W)
    unf = F.unfold(x.permute(0, 3, 1, 2).float(), 3, padding=1, stride=2)  # (B, C*9, L), index c*9 + tap
    unf = unf.reshape(B, C, 9, -1).permute(0, 3, 2, 1).reshape(B * (H // 2) * (W // 2), 9 * C)
    assert torch.equal(col.float(), unf)


def test_fp32_residual_stream_variants():
    """fp32 inputs (the engine's residual stream) for the norm / resample kernels + the raw bf16 copy."""
    from b200sd import ops
    from b200sd.packing import pack_conv3x3_f32
    torch.manual_seed(6)
    B, hw, C0, C1 = 2, 256, 1280, 640
    x0 = torch.randn(B * hw, C0, device=DEV) * 2 + 0.5
    x1 = torch.randn(B * hw, C1, device=DEV) - 1.0
    g, b = torch.randn(C0 + C1, device=DEV), torch.randn(C0 + C1, device=DEV)
    out = torch.empty(B * hw, C0 + C1, device=DEV, dtype=torch.bfloat16)
    raw = torch.empty_like(out)
    ops.groupnorm_silu(x0, x1, g, b, out, B, hw, 32, 1e-5, True, raw_out=raw)
    xc = torch.cat([x0, x1], 1)
    want = F.silu(F.group_norm(xc.reshape(B, hw, -1).permute(0, 2, 1), 32, g, b, 1e-5)).permute(0, 2, 1).reshape(B * hw, -1)
    _close(out, want, 1.0 / 128)
    assert torch.equal(raw, xc.bfloat16())

    x = torch.randn(2048, 640, device=DEV) * 3 + 1
    gg, bb = torch.randn(640, device=DEV), torch.randn(640, device=DEV)
    o = torch.empty(2048, 640, device=DEV, dtype=torch.bfloat16)
    ops.layernorm(x, gg, bb, o)
    _close(o, F.layer_norm(x, (640,), gg, bb, 1e-5), 1.0 / 128)

    Bn, H, W, C = 2, 16, 16, 640
    xf = torch.randn(Bn, H, W, C, device=DEV)
    up = torch.empty(Bn, 2 * H, 2 * W, C, device=DEV, dtype=torch.bfloat16)
    ops.upsample2x(xf.reshape(-1, C), up.reshape(-1, C), Bn, H, W)
    want = F.interpolate(xf.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).bfloat16()
    assert torch.equal(up, want)
    col = torch.empty(Bn * (H // 2) * (W // 2), 9 * C, device=DEV, dtype=torch.bfloat16)
    ops.im2col_s2(xf.reshape(-1, C), col, Bn, H, W)
    unf = F.unfold(xf.permute(0, 3, 1, 2), 3, padding=1, stride=2).reshape(Bn, C, 9, -1).permute(0, 3, 2, 1)
    assert torch.equal(col, unf.reshape(col.shape).bfloat16())

    xin = torch.randn(2, 4, 32, 32, device=DEV)
    w = torch.randn(320, 4, 3, 3, device=DEV) / 6
    bi = torch.randn(320, device=DEV)
    o32 = torch.empty(2 * 32 * 32, 320, device=DEV)
    ops.conv_in(xin, pack_conv3x3_f32(w), bi, o32)
    _close(o32, F.conv2d(xin, w, bi, padding=1).permute(0, 2, 3, 1).reshape(-1, 320), 1e-5)
