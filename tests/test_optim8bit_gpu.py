"""GPU parity of the block-wise 8-bit AdamW kernel (csrc/optim8bit.cu, b200sd_adamw8bit_step: the reference's default optimizer,
finetune_sd.py:300, 407-420) against oracle/adam8bit_ref.py -- byte / index work, so the bar is BIT-EXACT: moment codes, absmax
tables, fp32 moments of the small tensors, parameters and their bf16 copy -- and of FlatAdamW8bit / Trainer(optim_bits=8) on a
tiny UNet."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

from oracle import adam8bit_ref as A


def _random_modes(n_chunks, gen):
    """a flat-buffer-like layout: runs of 8-bit tensors, small fp32 tensors and frozen / padding chunks"""
    mode = torch.empty(n_chunks, dtype=torch.int32)
    k, small = 0, 0
    while k < n_chunks:
        kind = int(torch.randint(0, 10, (1,), generator=gen))
        run = int(torch.randint(1, 70, (1,), generator=gen)) if kind < 6 else int(torch.randint(1, 5, (1,), generator=gen))
        run = min(run, n_chunks - k)
        if kind < 6:
            mode[k:k + run] = A.MODE_8BIT
        elif kind < 9:
            mode[k:k + run] = torch.arange(small, small + run * 64, 64, dtype=torch.int32)
            small += run * 64
        else:
            mode[k:k + run] = A.MODE_SKIP
        k += run
    return mode


def _dense_code_books():
    """arbitrary sorted 256-entry code books with ~3 midpoints per lookup-table bin in their upper octave: the kernel must notice
    and take its generic eight-level search (the dynamic code books never do)"""
    lin = torch.linspace
    q1 = torch.cat([-lin(1, 0.5, 100), lin(-0.45, -0.01, 27), torch.zeros(1), lin(0.01, 0.45, 28), lin(0.5, 1, 100)])
    q2 = torch.cat([torch.zeros(1), lin(0.001, 0.45, 55), lin(0.5, 1, 200)])
    assert q1.numel() == 256 and q2.numel() == 256 and bool((q1[1:] > q1[:-1]).all()) and bool((q2[1:] > q2[:-1]).all())
    return q1.float(), q2.float()


def _medium_code_books():
    """up to two midpoints per lookup-table bin (150 codes over [0.5, 1]: spacing 0.0033 against bins of 0.0039), a first-moment
    book without a zero and with an odd number of negative codes: generic search again, different sign bookkeeping"""
    lin = torch.linspace
    q1 = torch.cat([-lin(1, 0.5, 75), lin(-0.4, -0.001, 16), lin(0.002, 0.4, 15), lin(0.5, 1, 150)])
    q2 = torch.cat([torch.zeros(1), lin(1e-5, 0.45, 105), lin(0.5, 1, 150)])
    assert q1.numel() == 256 and q2.numel() == 256 and bool((q1[1:] > q1[:-1]).all()) and bool((q2[1:] > q2[:-1]).all())
    return q1.float(), q2.float()


@pytest.mark.parametrize("n,with_modes,wd,dense", [(2048 * 5, False, 1e-2, False), (2048 * 37 + 64 * 7, True, 1e-2, False),
                                                   (64 * 3, True, 0.0, False), (2048 * 300 + 64, True, 0.1, False),
                                                   (2048 * 9 + 64 * 3, True, 1e-2, True), (2048 * 11 + 64, True, 0.0, "medium")])
def test_adamw8bit_kernel_is_bit_exact_against_the_oracle(n, with_modes, wd, dense):
    from b200sd import ops
    gen = torch.Generator().manual_seed(n)
    mode = _random_modes(n // 64, gen) if with_modes else None
    ref = A.AdamW8bitRef(n, mode, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    if dense:
        ref.qmap1, ref.qmap2 = _medium_code_books() if dense == "medium" else _dense_code_books()
    p_ref = torch.randn(n, generator=gen) * 0.1
    p = p_ref.clone().to(DEV)
    g_dev = torch.empty(n, device=DEV)
    wb = torch.full((n,), 7.0, dtype=torch.bfloat16, device=DEV)
    st1, st2 = torch.zeros(n, dtype=torch.uint8, device=DEV), torch.zeros(n, dtype=torch.uint8, device=DEV)
    am1, am2 = torch.zeros(ref.nblocks, device=DEV), torch.zeros(ref.nblocks, device=DEV)
    sm = torch.zeros(max(ref.small_m.numel(), 64), device=DEV)
    sv = torch.zeros_like(sm)
    q1, q2 = ref.qmap1.to(DEV), ref.qmap2.to(DEV)
    mode_dev = mode.to(DEV) if mode is not None else None
    scale = torch.logspace(-2, 0, n)[torch.randperm(n, generator=gen)]
    live = (mode.repeat_interleave(64) != A.MODE_SKIP) if mode is not None else torch.ones(n, dtype=torch.bool)
    for step in range(1, 5):
        g = torch.randn(n, generator=gen) * scale * 2.0           # the kernel averages a SUM over 2 ranks
        g[::97] = 0.0
        zero_grad = step != 2
        g_dev.copy_(g)
        want_wb = ref.step(p_ref, g, grad_scale=0.5, zero_grad=zero_grad)
        ops.adamw8bit_step(p, g_dev, st1, st2, am1, am2, q1, q2, mode_dev, sm, sv, wb, 3e-3, 0.9, 0.99, 1e-8, wd, step,
                           grad_scale=0.5, zero_grad=zero_grad)
        is8 = (mode.repeat_interleave(64) == A.MODE_8BIT) if mode is not None else live
        assert torch.equal(am1.cpu(), ref.absmax1) and torch.equal(am2.cpu(), ref.absmax2), f"absmax differs at step {step}"
        bad1 = int((st1.cpu()[is8] != ref.state1[is8]).sum())
        bad2 = int((st2.cpu()[is8] != ref.state2[is8]).sum())
        assert bad1 == 0 and bad2 == 0, f"step {step}: {bad1} first-moment / {bad2} second-moment codes differ of {int(is8.sum())}"
        assert torch.equal(p.cpu(), p_ref), f"parameters differ at step {step}: max {float((p.cpu() - p_ref).abs().max())}"
        assert torch.equal(g_dev.cpu(), g)                           # zeroed exactly where the oracle zeroed it
        assert torch.equal(wb.cpu()[live], want_wb[live]) and bool((wb.cpu()[~live] == 7.0).all())
        if ref.small_m.numel():
            assert torch.equal(sm.cpu()[:ref.small_m.numel()], ref.small_m) and torch.equal(sv.cpu()[:ref.small_v.numel()], ref.small_v)
    assert int(st1[~is8.to(DEV)].max() if bool((~is8).any()) else 0) == 0        # codes of non-8-bit elements are never written


def test_adamw8bit_rejects_bad_arguments():
    from b200sd import ops
    from b200sd._lib import B200SDError
    n = 2048
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=DEV)
    args = lambda n_: (z(n_), z(n_), z(n_, dt=torch.uint8), z(n_, dt=torch.uint8), z(1), z(1), z(256), z(256), None, None, None,
                       z(n_, dt=torch.bfloat16))
    with pytest.raises(B200SDError):
        ops.adamw8bit_step(*args(n + 8), 1e-3, 0.9, 0.999, 1e-8, 0.0, 1)             # not a multiple of 64
    with pytest.raises(B200SDError):
        ops.adamw8bit_step(*args(n), 1e-3, 0.9, 0.999, 1e-8, 0.0, 0)                 # step counts from 1
    a = list(args(n))
    a[8] = torch.full((n // 64,), -1, dtype=torch.int32, device=DEV)                 # a chunk table without the fp32 side buffers
    with pytest.raises(B200SDError):
        ops.adamw8bit_step(*a, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1)


def test_trainer_with_8bit_optimizer_state_trains_the_tiny_unet():
    """Trainer(optim_bits=8) == the reference's default `bnb.optim.AdamW8bit(..., min_8bit_size=16384)`: the fixed batch is fitted
    about as fast as with fp32 moments, the optimizer state is ~2 B per parameter, frozen parameters stay put, and the
    torch.optim.Optimizer surface (lr scheduler, state_dict) works."""
    from b200sd.schedulers import DDPMScheduler
    from b200sd.trainer import FlatAdamW8bit, Trainer
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import TINY_OVERRIDES
    torch.backends.cuda.matmul.allow_tf32 = False
    sched = DDPMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000)
    g = torch.Generator().manual_seed(5)
    x, noise = torch.randn(4, 4, 32, 32, generator=g).to(DEV), torch.randn(4, 4, 32, 32, generator=g).to(DEV)
    ctx, t = torch.randn(4, 77, 64, generator=g).to(DEV), torch.randint(0, 1000, (4,), generator=g).to(DEV)
    losses = {}
    for bits in (32, 8):
        torch.manual_seed(0)
        unet = UNet2DConditionModel(**TINY_OVERRIDES).to(DEV)
        unet.conv_in.weight.requires_grad_(False)
        frozen = unet.conv_in.weight.detach().clone()
        tr = Trainer(unet, sched, lr=2e-4, weight_decay=1e-2, optim_bits=bits)
        opt = tr.optimizer()
        lr_sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=40, eta_min=1e-6)
        ls = []
        for _ in range(12):
            ls.append(float(tr.train_step(x, noise, t, ctx)))
            lr_sched.step()
        losses[bits] = ls
        assert torch.equal(unet.conv_in.weight.detach(), frozen) and opt.steps == 12 and opt.lr < 2e-4
        if bits == 8:
            assert isinstance(opt, FlatAdamW8bit)
            n = opt.flat.master.numel()
            assert opt.state_bytes() < 2.5 * n, (opt.state_bytes(), n)             # 2 B codes + absmax + the small tensors' fp32 (3.8 % of the tiny net)
            st = opt.state[opt.flat.master]
            assert int(st["state1"].max()) > 0 and float(st["absmax2"].max()) > 0 and float(st["exp_avg"].abs().max()) > 0
            sd = opt.state_dict()
            opt2 = FlatAdamW8bit(opt.flat, lr=1.0)
            opt2.load_state_dict(sd)
            st2 = opt2.state[opt2.flat.master]
            assert opt2.steps == 12 and st2["state1"].dtype == torch.uint8 and torch.equal(st2["state1"], st["state1"])
            assert torch.equal(st2["absmax1"], st["absmax1"]) and opt2.lr == opt.lr
            # the weights the inference engine sees are the updated ones
            unet.eval()
            with torch.no_grad():
                a = unet(x, t, ctx).sample
            unet.train()
            b = unet(x, t, ctx).sample.detach()
            assert float((a - b).abs().max() / b.abs().max()) <= 2e-2
    assert losses[8][-1] < 0.7 * losses[8][0], losses[8]
    assert abs(losses[8][-1] - losses[32][-1]) <= 0.15 * losses[32][0], (losses[8][-1], losses[32][-1])
