"""CPU self-checks of the block-wise 8-bit AdamW oracle (oracle/adam8bit_ref.py: bitsandbytes 0.35.4 `AdamW8bit` restated; PARITY
UNPINNED -- bitsandbytes is not installable here and the reference holds no vector for it).  What can be checked without the
library: the structure of the dynamic code books, the quantisation error bounds that structure implies, the fp32 (small-tensor)
branch against the closed-form update, the sign rule, and that training with 8-bit moments follows fp32 torch.optim.AdamW."""
import torch

from oracle import adam8bit_ref as A


def test_dynamic_code_books_structure():
    s, u = A.create_dynamic_map(True), A.create_dynamic_map(False)
    for q in (s, u):
        assert q.shape == (256,) and q.dtype == torch.float32 and bool((q[1:] > q[:-1]).all())
        assert float(q[-1]) == 1.0 and bool((q == 0).sum() == 1)
    # signed: 127 negative values mirror the 127 positive fractions (1.0 has no mirror), 0 sits at code 127
    assert float(s[127]) == 0.0 and torch.equal(s[:127], -s[128:255].flip(0))
    assert float(u[0]) == 0.0 and float(u.min()) == 0.0
    # decade i (values in (10^(i-7), 10^(i-6))) holds 2^i positive fractions (signed) / 2^(i+1) (unsigned)
    for i in range(7):
        lo, hi = 10.0 ** (i - 7), 10.0 ** (i - 6)
        assert int(((s > lo) & (s < hi)).sum()) == 2 ** i
        assert int(((u > lo) & (u < hi)).sum()) == 2 ** (i + 1)
    # the product keeps its own copy of the construction (it may not import the oracle): same tables
    from b200sd.trainer import create_dynamic_map
    assert torch.equal(create_dynamic_map(True), s) and torch.equal(create_dynamic_map(False), u)


def test_nearest_code_and_blockwise_round_trip_error():
    torch.manual_seed(0)
    q = A.create_dynamic_map(True)
    x = torch.rand(10000) * 2 - 1
    codes = A.quantize_nearest(x, q).long()
    brute = (x[:, None] - q[None, :]).abs().argmin(dim=1)
    assert bool(((x - q[codes]).abs() <= (x - q[brute]).abs() + 1e-7).all())       # nearest (ties may pick either neighbour)
    assert torch.equal(A.quantize_nearest(q, q).long(), torch.arange(256))           # every code is a fixed point
    # block-wise: the top decade has 64 fractions over (0.1, 1): |error| <= 0.9 / 64 / 2 of the block's absmax
    v = torch.randn(5 * A.BLOCK + 100) * torch.logspace(-3, 1, 5 * A.BLOCK + 100)
    c, absmax = A.quantize_blockwise(v, q)
    assert absmax.shape == (6,) and c.dtype == torch.uint8
    back = A.dequantize_blockwise(c, absmax, q)
    scale = absmax[torch.arange(v.numel()) // A.BLOCK]
    big = v.abs() >= 0.1 * scale
    assert float(((back - v).abs() / scale)[big].max()) <= 0.9 / 64 / 2 + 1e-6
    assert float(((back - v).abs() / scale).max()) <= 0.9 / 64 / 2 + 1e-6           # smaller decades are finer still
    assert bool((torch.sign(back[v.abs() > 1e-5 * scale]) == torch.sign(v[v.abs() > 1e-5 * scale])).all())


def _fp32_adamw_bnb_order(p, g, m, v, lr, b1, b2, eps, wd, step):
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    c1, c2 = 1 - b1 ** step, (1 - b2 ** step) ** 0.5
    p.add_(m / (v.sqrt() + c2 * eps), alpha=-lr * c2 / c1)
    p.mul_(1 - lr * wd)


def test_small_tensor_branch_is_plain_fp32_adam_and_skipped_chunks_are_untouched():
    torch.manual_seed(1)
    n = 64 * 6
    mode = torch.tensor([A.MODE_SKIP, 0, 64, A.MODE_SKIP, 128, A.MODE_SKIP], dtype=torch.int32)
    opt = A.AdamW8bitRef(n, mode, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.1)
    p = torch.randn(n)
    p0 = p.clone()
    rp, rm, rv = p.clone(), torch.zeros(n), torch.zeros(n)
    live = mode.repeat_interleave(64) >= 0
    for step in range(1, 4):
        g = torch.randn(n)
        g_keep = g.clone()
        _fp32_adamw_bnb_order(rp, g, rm, rv, 1e-2, 0.9, 0.99, 1e-8, 0.1, step)
        wb = opt.step(p, g)
        assert torch.equal(g[~live], g_keep[~live]) and float(g[live].abs().max()) == 0.0     # zero_grad only where it stepped
        torch.testing.assert_close(p[live], rp[live], rtol=1e-5, atol=1e-7)
        assert torch.equal(p[~live], p0[~live]) and wb.dtype == torch.bfloat16
    m, v = opt.moments()
    torch.testing.assert_close(m[live], rm[live], rtol=1e-5, atol=1e-8)
    assert float(opt.absmax1.abs().max()) == 0.0 and int(opt.state1.max()) == 0               # no 8-bit element anywhere


def test_first_moment_keeps_its_sign():
    """bitsandbytes: "make sure state1 term has still the same sign after quantization" -- a tiny negative moment next to a
    large one must not come back as +0."""
    opt = A.AdamW8bitRef(A.BLOCK, lr=1e-3, betas=(0.0, 0.0), weight_decay=0.0)
    g = torch.zeros(A.BLOCK)
    g[0], g[1], g[2] = 1.0, -1e-9, 1e-9
    opt.step(torch.zeros(A.BLOCK), g.clone())
    m, _ = opt.moments()
    assert float(m[0]) == 1.0 and float(m[1]) < 0 and int(opt.state1[1]) == 126 and int(opt.state1[2]) == 127


def test_8bit_moments_follow_fp32_adamw_over_a_training_like_sequence():
    torch.manual_seed(2)
    n = 4 * A.BLOCK
    p8 = torch.randn(n) * 0.05
    ref = p8.clone().requires_grad_(True)
    p_start = p8.clone()
    opt = A.AdamW8bitRef(n, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    ropt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    # gradient magnitudes over two decades inside every block.  (Beyond ~3.4 decades the scheme itself breaks down, in bitsandbytes
    # as here: a second moment below 1.6e-7 of its block's maximum rounds to the code 0 and m / (sqrt(0) + eps) explodes.)
    scale = torch.logspace(-2, 0, n)[torch.randperm(n)]
    drift = torch.randn(n)
    for _ in range(30):
        g = (drift + 0.5 * torch.randn(n)) * scale
        ref.grad = g.clone()
        ropt.step()
        opt.step(p8, g.clone())
    moved8, moved32 = p8 - p_start, ref.detach() - p_start
    cos = float(torch.nn.functional.cosine_similarity(moved8, moved32, dim=0))
    rel = float((moved8 - moved32).norm() / moved32.norm())
    assert cos >= 0.995 and rel <= 0.1, (cos, rel)            # measured 0.9986 / 0.053
    assert opt.state1.numel() + opt.state2.numel() + 4 * (opt.absmax1.numel() + opt.absmax2.numel()) < 2.01 * n   # ~2 B / parameter


def test_committed_regression_vectors():
    """tests/golden/adam8bit_oracle_kat.json (generated by the oracle itself: a regression guard, not a pin)"""
    import importlib.util
    import json
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_adam8bit_golden", os.path.join(here, "make_adam8bit_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(here, "adam8bit_oracle_kat.json")))
    assert mod.run() == want
