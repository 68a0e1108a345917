"""GPU parity of the non-GEMM backward kernels against torch autograd on the same inputs:
attention backward, GroupNorm(+SiLU) / LayerNorm backward, GEGLU, resampler transposes, the 4-channel end
convs, gradient prep.  The autograd reference runs in fp32 on the bf16-rounded operands the kernels see;
tolerances are bf16 output rounding (2^-7 of the tensor's max) or 2e-4 for fp32 outputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(got, want, rel, what=""):
    got, want = got.float(), want.float()
    scale = float(want.abs().max()) + 1e-9
    err = float((got - want).abs().max())
    assert err <= rel * scale, f"{what}: max abs err {err:.4g} vs scale {scale:.4g} (rel {err / scale:.3g})"


def _setup(seed=0):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(seed)


@pytest.mark.parametrize("B,H,Sq,Skv,d", [(2, 8, 1024, 1024, 40), (1, 8, 4096, 4096, 40), (2, 8, 256, 256, 80),
                                          (2, 8, 64, 64, 160), (2, 8, 1024, 77, 40), (2, 8, 96, 77, 160),
                                          (2, 2, 256, 256, 32), (1, 2, 16, 77, 64), (1, 8, 384, 384, 80)])
def test_attention_bwd(B, H, Sq, Skv, d):
    from b200sd import ops
    _setup(Sq + d)
    C = H * d
    q = torch.randn(B * Sq, C, device=DEV).bfloat16()
    k = torch.randn(B * Skv, C, device=DEV).bfloat16()
    v = torch.randn(B * Skv, C, device=DEV).bfloat16()
    do = torch.randn(B * Sq, C, device=DEV).bfloat16()
    scale = d ** -0.5
    out = torch.empty_like(q)
    lse = torch.empty(B, H, Sq, device=DEV)
    ops.attention_lse(q, k, v, out, lse, B, H, Sq, Skv, d, scale)
    qf, kf, vf = (t.float().reshape(B, -1, H, d).transpose(1, 2).requires_grad_(True) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * scale
    want_lse = torch.logsumexp(s, -1) * 1.4426950408889634
    o = torch.softmax(s, -1) @ vf
    _close(out, o.transpose(1, 2).reshape(B * Sq, C), 1.0 / 64, "out")
    _close(lse, want_lse, 1e-3, "lse")
    o.backward(do.float().reshape(B, Sq, H, d).transpose(1, 2))
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops.attention_bwd(q, k, v, out, do, lse, dq, dk, dv, B, H, Sq, Skv, d, scale)
    for name, got, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        _close(got, ref.transpose(1, 2).reshape(got.shape), 1.0 / 50, name)


def test_attention_bwd_fused_qkv_slices():
    """q/k/v and dq/dk/dv as column slices of [M, 3C] buffers (the self-attention layout of the engine)."""
    from b200sd import ops
    _setup(3)
    B, H, S, d = 2, 8, 256, 40
    C = H * d
    qkv = torch.randn(B * S, 3 * C, device=DEV).bfloat16()
    do = torch.randn(B * S, C, device=DEV).bfloat16()
    out = torch.empty(B * S, C, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device=DEV)
    ops.attention_lse(qkv, qkv, qkv, out, lse, B, H, S, S, d, d ** -0.5, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, k_off=C, v_off=2 * C)
    dqkv = torch.zeros_like(qkv)
    ops.attention_bwd(qkv, qkv, qkv, out, do, lse, dqkv, dqkv, dqkv, B, H, S, S, d, d ** -0.5, ldq=3 * C, ldk=3 * C, ldv=3 * C,
                      lddq=3 * C, lddk=3 * C, lddv=3 * C, k_off=C, v_off=2 * C, dk_off=C, dv_off=2 * C)
    x = qkv.float().requires_grad_(True)
    qf, kf, vf = (t.reshape(B, S, H, d).transpose(1, 2) for t in x.split(C, dim=1))
    o = torch.softmax((qf @ kf.transpose(-1, -2)) * d ** -0.5, -1) @ vf
    o.backward(do.float().reshape(B, S, H, d).transpose(1, 2))
    _close(dqkv, x.grad, 1.0 / 50, "dqkv")


@pytest.mark.parametrize("B,hw,C0,C1,silu,in_f32,out_f32", [(2, 4096, 320, 0, True, True, True), (2, 1024, 640, 320, True, True, True),
                                                            (3, 256, 1280, 1280, True, True, False), (2, 64, 1280, 0, False, True, True),
                                                            (2, 1024, 320, 0, True, False, False), (2, 16, 64, 0, True, True, True),
                                                            (1, 6144, 320, 0, False, True, True), (2, 256, 1280, 640, True, True, True)])
def test_groupnorm_silu_bwd(B, hw, C0, C1, silu, in_f32, out_f32):
    from b200sd import ops
    _setup(hw + C0)
    C = C0 + C1
    dt = torch.float32 if in_f32 else torch.bfloat16
    x0 = (torch.randn(B * hw, C0, device=DEV) * 1.5 + 0.3).to(dt)
    x1 = (torch.randn(B * hw, C1, device=DEV) * 0.7).to(dt) if C1 else None
    gamma = torch.randn(C, device=DEV) * 0.5 + 1
    beta = torch.randn(C, device=DEV) * 0.2
    dy = torch.randn(B * hw, C, device=DEV).bfloat16()
    add = torch.randn(B * hw, C, device=DEV)
    xin = torch.cat([x0, x1], 1) if C1 else x0
    xr = xin.float().reshape(B, hw, C).permute(0, 2, 1).requires_grad_(True)
    g_, b_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xr, 32, g_, b_, 1e-5)
    if silu:
        y = F.silu(y)
    y.backward(dy.float().reshape(B, hw, C).permute(0, 2, 1))
    want = xr.grad.permute(0, 2, 1).reshape(B * hw, C) + add
    odt = torch.float32 if out_f32 else torch.bfloat16
    # forward kernel (also writes the (mean, rstd) the fast backward path consumes)
    t_fwd = torch.empty(B * hw, C, device=DEV, dtype=torch.bfloat16)
    st = torch.empty(B, 32, 2, device=DEV)
    ops.groupnorm_silu(x0, x1, gamma, beta, t_fwd, B, hw, 32, 1e-5, silu, stats_out=st)
    _close(t_fwd, y.detach().permute(0, 2, 1).reshape(B * hw, C), 1.0 / 64, "forward")
    xg = xr.detach().reshape(B, 32, -1)
    _close(st[..., 0], xg.mean(-1), 1e-4, "saved mean")
    _close(st[..., 1], (xg.var(-1, unbiased=False) + 1e-5).rsqrt(), 1e-4, "saved rstd")
    for mean_rstd in (None, st):      # recompute-statistics path and saved-statistics fast path
        prev0 = torch.randn(B * hw, C0, device=DEV).to(odt)
        out0 = prev0.clone()
        out1 = torch.empty(B * hw, C1, device=DEV, dtype=odt) if C1 else None
        dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        ops.groupnorm_silu_bwd(x0, x1, gamma, beta, dy, out0, out1, B, hw, add_src=add, acc0=True, acc1=False, dgamma=dg,
                               dbeta=db, eps=1e-5, silu=silu, mean_rstd=mean_rstd)
        tol = 2e-4 if out_f32 else 1.0 / 100
        tag = "fast" if mean_rstd is not None else "recompute"
        _close(out0, want[:, :C0] + prev0.float(), tol, f"dx0 (accumulated, {tag})")
        if C1:
            _close(out1, want[:, C0:], tol, f"dx1 ({tag})")
        _close(dg, g_.grad, 2e-4, f"dgamma ({tag})")
        _close(db, b_.grad, 2e-4, f"dbeta ({tag})")


@pytest.mark.parametrize("rows,C,in_f32", [(8192, 320, True), (2048, 640, True), (512, 1280, True), (100, 64, True), (1024, 320, False)])
def test_layernorm_bwd(rows, C, in_f32):
    from b200sd import ops
    _setup(rows + C)
    x = (torch.randn(rows, C, device=DEV) * 2 + 0.5).to(torch.float32 if in_f32 else torch.bfloat16)
    gamma = torch.randn(C, device=DEV) * 0.5 + 1
    dy = torch.randn(rows, C, device=DEV).bfloat16()
    xr = x.float().requires_grad_(True)
    g_ = gamma.clone().requires_grad_(True)
    b_ = torch.zeros(C, device=DEV, requires_grad=True)
    F.layer_norm(xr, (C,), g_, b_, 1e-5).backward(dy.float())
    dres0 = torch.randn(rows, C, device=DEV)
    dres = dres0.clone()
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.layernorm_bwd(x, gamma, dy, dres, dg, db)
    _close(dres, dres0 + xr.grad, 2e-4, "dres")
    _close(dg, g_.grad, 2e-4, "dgamma")
    _close(db, b_.grad, 2e-4, "dbeta")


def test_geglu_fwd_bwd():
    from b200sd import ops
    _setup(5)
    rows, Ch = 1000, 1280
    u = torch.randn(rows, 2 * Ch, device=DEV).bfloat16()
    dff = torch.randn(rows, Ch, device=DEV).bfloat16()
    ur = u.float().requires_grad_(True)
    val, gate = ur.chunk(2, dim=1)
    y = val * F.gelu(gate)
    y.backward(dff.float())
    out = torch.empty(rows, Ch, device=DEV, dtype=torch.bfloat16)
    du = torch.empty_like(u)
    ops.geglu_fwd(u, out)
    ops.geglu_bwd(u, dff, du)
    _close(out, y, 1.0 / 128, "geglu fwd")
    _close(du, ur.grad, 1.0 / 128, "geglu bwd")


def test_resampler_transposes():
    from b200sd import ops
    _setup(6)
    B, H, W, C = 2, 16, 8, 64
    # upsample
    dy = torch.randn(B, 2 * H, 2 * W, C, device=DEV).bfloat16()
    xr = torch.zeros(B, C, H, W, device=DEV, requires_grad=True)
    F.interpolate(xr, scale_factor=2.0, mode="nearest").backward(dy.float().permute(0, 3, 1, 2))
    prev = torch.randn(B * H * W, C, device=DEV)
    dx = prev.clone()
    ops.upsample2x_bwd(dy.reshape(-1, C), dx, B, H, W, accumulate=True)
    _close(dx, prev + xr.grad.permute(0, 2, 3, 1).reshape(-1, C), 1e-5, "upsample bwd")
    # stride-2 conv: col2im is the transpose of im2col_s2
    x = torch.randn(B * H * W, C, device=DEV)
    col = torch.empty(B * (H // 2) * (W // 2), 9 * C, device=DEV, dtype=torch.bfloat16)
    ops.im2col_s2(x, col, B, H, W)
    dcol = torch.randn_like(col.float()).bfloat16()
    dx2 = torch.empty(B * H * W, C, device=DEV)
    ops.col2im_s2(dcol, dx2, B, H, W)
    # <im2col(x), dcol> == <x, col2im(dcol)>
    xb = x.bfloat16().float()
    lhs = float((col.float() * dcol.float()).sum())
    rhs = float((xb * dx2).sum())
    assert abs(lhs - rhs) <= 2e-3 * (abs(lhs) + 1), (lhs, rhs)
    # and against autograd of the strided conv itself
    w = torch.randn(C, C, 3, 3, device=DEV) / 24
    xr2 = xb.reshape(B, H, W, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
    yy = F.conv2d(xr2, w, stride=2, padding=1)
    gy = torch.randn_like(yy)
    yy.backward(gy)
    wp = w.permute(0, 2, 3, 1).reshape(C, 9 * C)
    dcol_ref = (gy.permute(0, 2, 3, 1).reshape(-1, C) @ wp).bfloat16()
    ops.col2im_s2(dcol_ref, dx2, B, H, W)
    _close(dx2, xr2.grad.permute(0, 2, 3, 1).reshape(-1, C), 1.0 / 100, "downsample dgrad")


def test_end_convs_bwd():
    from b200sd import ops, packing
    _setup(7)
    B, H, W, Cw = 2, 16, 16, 320
    # conv_out: 320 -> 4
    t = torch.randn(B * H * W, Cw, device=DEV).bfloat16()
    w = torch.randn(4, Cw, 3, 3, device=DEV) / 50
    dout = torch.randn(B, 4, H, W, device=DEV)
    tr = t.float().reshape(B, H, W, Cw).permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = torch.zeros(4, device=DEV, requires_grad=True)
    F.conv2d(tr, wr, br, padding=1).backward(dout)
    dx = torch.empty_like(t)
    dw = torch.zeros(4, 9 * Cw, device=DEV)
    dbias = torch.zeros(4, device=DEV)
    ops.conv_out_bwd(dout, t, packing.pack_conv3x3_f32(w), dx, dw, dbias)
    _close(dx, tr.grad.permute(0, 2, 3, 1).reshape(-1, Cw), 1.0 / 128, "conv_out dgrad")
    _close(dw, wr.grad.permute(0, 2, 3, 1).reshape(4, -1), 2e-4, "conv_out wgrad")
    _close(dbias, br.grad, 2e-4, "conv_out dbias")
    # conv_in: 4 -> 320
    x = torch.randn(B, 4, H, W, device=DEV)
    dy = torch.randn(B * H * W, Cw, device=DEV)
    w_in = (torch.randn(Cw, 4, 3, 3, device=DEV) / 6).requires_grad_(True)
    F.conv2d(x, w_in, padding=1).backward(dy.reshape(B, H, W, Cw).permute(0, 3, 1, 2))
    dwi = torch.zeros(Cw, 36, device=DEV)
    ops.conv_in_wgrad(dy, x, dwi)
    _close(dwi, w_in.grad.permute(0, 2, 3, 1).reshape(Cw, 36), 2e-4, "conv_in wgrad")


def test_grad_prep_and_small_helpers():
    from b200sd import ops
    _setup(8)
    B, hw, N = 4, 256, 320
    g = torch.randn(B * hw, N, device=DEV)
    gb = torch.empty(B * hw, N, device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(N, device=DEV)
    ops.grad_prep(g, gb, cs)
    assert torch.equal(gb, g.bfloat16())
    _close(cs, g.sum(0), 1e-5, "colsum")
    per = torch.zeros(B, 1000, device=DEV)
    ops.grad_prep(gb, None, per.view(-1)[40:], rows_per_image=hw, ldcs=1000)   # per-image sums into a column window
    _close(per[:, 40:40 + N], gb.float().reshape(B, hw, N).sum(1), 1e-5, "per-image colsum")
    assert float(per[:, :40].abs().max()) == 0 and float(per[:, 40 + N:].abs().max()) == 0
    x = torch.randn(8, 1280, device=DEV)
    xb = torch.empty(8, 1280, device=DEV, dtype=torch.bfloat16)
    ops.cast_act(x, xb, silu=True)
    _close(xb, F.silu(x), 1.0 / 128, "cast_act")
    xr = x.clone().requires_grad_(True)
    gr = torch.randn_like(x)
    F.silu(xr).backward(gr)
    g2 = gr.clone()
    ops.silu_bwd_mul(x, g2)
    _close(g2, xr.grad, 1e-5, "silu_bwd_mul")
