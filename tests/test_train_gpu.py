"""GPU parity of the UNet training path (BASELINE configs 3 and 4): the gradients of
`mse(unet(noisy, t, ctx).sample, noise)` w.r.t. all 686 parameters (and w.r.t. the text context) against
fp32 torch autograd through the oracle on identical random-init weights and inputs.

Tolerance: the backward runs bf16 tensor-core operands with fp32 accumulation, like the forward, whose bar is
1e-2 max-relative on the noise prediction.  Gradients cross the network twice, so the per-tensor bars are
max|g - ref| / max|ref| <= 8e-2 and cosine >= 0.998, and the whole flattened gradient must reach cosine
>= 0.9995 (measured on B200: worst tensor 5e-2 / 0.9993 on the tiny net; torch's own bf16-autocast backward of
the oracle is printed beside it as the library noise level).  A wrong kernel shows up as O(1), not as 1e-2."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _pair(overrides):
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import make_oracle_unet
    _setup()
    oracle = make_oracle_unet(seed=0, **overrides).to(DEV)
    ours = UNet2DConditionModel(**overrides)
    ours.load_state_dict(oracle.state_dict(), strict=True)
    return oracle, ours.to(DEV).train()


def _inputs(N, h, w, ctx_dim, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, 4, h, w, generator=g).to(DEV)
    noise = torch.randn(N, 4, h, w, generator=g).to(DEV)
    ctx = torch.randn(N, 77, ctx_dim, generator=g).to(DEV)
    t = torch.randint(0, 1000, (N,), generator=g).to(DEV)
    return x, noise, ctx, t


def _compare(named_ref, named_got, bar_rel=8e-2, bar_cos=0.998):
    worst, flat_r, flat_g, bad = 0.0, [], [], []
    for name, ref in named_ref.items():
        got = named_got[name]
        assert got is not None, f"{name}: no gradient"
        assert got.shape == ref.shape, name
        r, g = ref.float().flatten(), got.float().flatten()
        scale = float(r.abs().max())
        rel = float((g - r).abs().max()) / (scale + 1e-20)
        cos = float(F.cosine_similarity(g, r, dim=0))
        flat_r.append(r)
        flat_g.append(g)
        worst = max(worst, rel)
        if rel > bar_rel or cos < bar_cos:
            bad.append((name, rel, cos))
    cos_all = float(F.cosine_similarity(torch.cat(flat_g), torch.cat(flat_r), dim=0))
    return worst, cos_all, bad


def _ref_grads(oracle, x, noise, ctx, t, ctx_grad=False):
    oracle.zero_grad(set_to_none=True)
    c = ctx.clone().requires_grad_(ctx_grad)
    loss = F.mse_loss(oracle(x, t, c).sample, noise)
    loss.backward()
    grads = {n: p.grad.clone() for n, p in oracle.named_parameters() if p.grad is not None}
    return float(loss), grads, (c.grad.clone() if ctx_grad else None)


def test_tiny_unet_all_gradients_and_context_gradient():
    from b200sd import ops
    from oracle.unet_ref import TINY_OVERRIDES
    oracle, ours = _pair(TINY_OVERRIDES)
    x, noise, ctx, t = _inputs(2, 32, 32, 64, 0)
    ref_loss, ref, ref_dctx = _ref_grads(oracle, x, noise, ctx, t, ctx_grad=True)
    c = ctx.clone().requires_grad_(True)
    loss = ops.mse_loss(ours(x, t, c).sample, noise)
    loss.backward()
    assert abs(float(loss) - ref_loss) <= 2e-2 * ref_loss
    got = {n: p.grad for n, p in ours.named_parameters()}
    worst, cos_all, bad = _compare(ref, got)
    print(f"tiny grads: worst per-tensor max-rel {worst:.4g}, global cosine {cos_all:.6f}")
    assert not bad, f"{len(bad)} tensors out of tolerance, e.g. {bad[:5]}"
    assert cos_all >= 0.9995, cos_all
    rel = float((c.grad - ref_dctx).abs().max() / ref_dctx.abs().max())
    assert rel <= 4e-2, f"context gradient max-rel {rel}"
    # library noise level: torch's bf16-autocast backward of the same oracle
    oracle.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lb = F.mse_loss(oracle(x, t, ctx).sample.float(), noise)
    lb.backward()
    w2, c2, _ = _compare(ref, {n: p.grad for n, p in oracle.named_parameters()})
    print(f"torch bf16 autocast backward of the oracle: worst per-tensor max-rel {w2:.4g}, global cosine {c2:.6f}")


def test_direct_gradients_accumulate_in_the_flat_buffer():
    from b200sd import ops
    from oracle.unet_ref import TINY_OVERRIDES
    oracle, ours = _pair(TINY_OVERRIDES)
    ours.enable_direct_gradients()
    x, noise, ctx, t = _inputs(2, 32, 32, 64, 1)
    _, ref, _ = _ref_grads(oracle, x, noise, ctx, t)
    for _ in range(2):   # two micro-steps without zero_grad: gradient accumulation (finetune_sd.py:454-458)
        ops.mse_loss(ours(x, t, ctx).sample, noise).backward()
    flat = ours.flat_gradients()
    p = ours.down_blocks[0].resnets[0].conv1.weight
    assert p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr(), "param.grad must alias the flat buffer"
    got = {n: q.grad / 2 for n, q in ours.named_parameters()}
    _, cos_all, bad = _compare(ref, got)
    assert not bad and cos_all >= 0.9995, (cos_all, bad[:5])
    # optimizer.zero_grad() sets grads to None: the next backward starts from zero again
    for q in ours.parameters():
        q.grad = None
    ops.mse_loss(ours(x, t, ctx).sample, noise).backward()
    _, cos_all, bad = _compare(ref, {n: q.grad for n, q in ours.named_parameters()})
    assert not bad and cos_all >= 0.9995, (cos_all, bad[:5])


def test_frozen_unet_context_gradient_only():
    """BASELINE config 4 (train_text_encoder): UNet frozen, gradient flows to encoder_hidden_states only."""
    from b200sd import ops
    from oracle.unet_ref import TINY_OVERRIDES
    oracle, ours = _pair(TINY_OVERRIDES)
    ours.requires_grad_(False).eval()
    x, noise, ctx, t = _inputs(2, 32, 32, 64, 2)
    _, _, ref_dctx = _ref_grads(oracle, x, noise, ctx, t, ctx_grad=True)
    c = ctx.clone().requires_grad_(True)
    ops.mse_loss(ours(x, t, c).sample, noise).backward()
    assert all(p.grad is None for p in ours.parameters())
    rel = float((c.grad - ref_dctx).abs().max() / ref_dctx.abs().max())
    assert rel <= 4e-2, f"context gradient max-rel {rel}"


def test_sd15_unet_gradients_batch2():
    """Full SD v1.5 UNet (859.5 M parameters), batch 2 at 64x64 latents."""
    from b200sd import ops
    oracle, ours = _pair({})
    x, noise, ctx, t = _inputs(2, 64, 64, 768, 3)
    ref_loss, ref, ref_dctx = _ref_grads(oracle, x, noise, ctx, t, ctx_grad=True)
    oracle.zero_grad(set_to_none=True)
    c = ctx.clone().requires_grad_(True)
    loss = ops.mse_loss(ours(x, t, c).sample, noise)
    loss.backward()
    assert abs(float(loss) - ref_loss) <= 2e-2 * ref_loss
    worst, cos_all, bad = _compare(ref, {n: p.grad for n, p in ours.named_parameters()})
    print(f"sd15 grads: worst per-tensor max-rel {worst:.4g}, global cosine {cos_all:.6f}, out-of-tolerance {len(bad)}")
    assert not bad, f"{len(bad)} tensors out of tolerance, e.g. {bad[:5]}"
    assert cos_all >= 0.9995, cos_all
    rel = float((c.grad - ref_dctx).abs().max() / ref_dctx.abs().max())
    assert rel <= 4e-2, f"context gradient max-rel {rel}"


def test_flat_adamw_matches_torch_adamw():
    """One fused kernel over the flat buffers == torch.optim.AdamW on the individual parameters (fp32, 3 steps)."""
    from b200sd.train import FlatParams
    from b200sd.trainer import FlatAdamW
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import TINY_OVERRIDES
    _setup()
    torch.manual_seed(0)
    ours = UNet2DConditionModel(**TINY_OVERRIDES).to(DEV)
    ref_params = [p.detach().clone().requires_grad_(True) for p in ours.parameters()]
    flat = FlatParams(ours, torch.device(DEV))
    flat.attach_grads()
    opt = FlatAdamW(flat, lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.1)
    ref = torch.optim.AdamW(ref_params, lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.1)
    for step in range(3):
        for p, q in zip(ours.parameters(), ref_params):
            g = torch.randn_like(q)
            q.grad = g
            p.grad.copy_(g * 2)             # the fused step averages a SUM over 2 ranks
        opt.step(grad_scale=0.5, zero_grad=True)
        ref.step()
    assert float(flat.grad.abs().max()) == 0.0
    for (n, p), q in zip(ours.named_parameters(), ref_params):
        err = float((p.detach() - q.detach()).abs().max())
        assert err <= 2e-6 * (1 + float(q.abs().max())), (n, err)
    assert torch.equal(flat.wb, flat.master.bfloat16())


def test_flat_adamw_under_the_reference_cosine_lr_schedule():
    """finetune_sd.py:421-422, 577: torch's CosineAnnealingLR(optimizer, T_max, eta_min=1e-6) hangs on FlatAdamW (a
    torch.optim.Optimizer) and the fused kernel follows torch.optim.AdamW under the same schedule; optimizer.zero_grad() and the
    state_dict round trip of the reference's surface work on it."""
    from b200sd.train import FlatParams
    from b200sd.trainer import FlatAdamW
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import TINY_OVERRIDES
    _setup()
    torch.manual_seed(0)
    ours = UNet2DConditionModel(**TINY_OVERRIDES).to(DEV)
    ref_params = [p.detach().clone().requires_grad_(True) for p in ours.parameters()]
    flat = FlatParams(ours, torch.device(DEV))
    flat.attach_grads()
    opt = FlatAdamW(flat, lr=1e-3, weight_decay=1e-2)
    ref = torch.optim.AdamW(ref_params, lr=1e-3, weight_decay=1e-2)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=5, eta_min=1e-6)
    rsched = torch.optim.lr_scheduler.CosineAnnealingLR(ref, T_max=5, eta_min=1e-6)
    for step in range(4):
        for p, q in zip(ours.parameters(), ref_params):
            g = torch.randn_like(q)
            q.grad = g
            p.grad.copy_(g)
        opt.step()
        opt.zero_grad()
        ref.step()
        sched.step()
        rsched.step()
        assert opt.lr == ref.param_groups[0]["lr"]
    assert opt.steps == 4 and opt.lr < 2e-4
    for (n, p), q in zip(ours.named_parameters(), ref_params):
        err = float((p.detach() - q.detach()).abs().max())
        assert err <= 2e-6 * (1 + float(q.abs().max())), (n, err)
    sd = opt.state_dict()
    opt2 = FlatAdamW(flat, lr=1.0)
    opt2.load_state_dict(sd)
    assert opt2.lr == opt.lr and opt2.steps == 4 and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq)
    assert opt2.exp_avg.device == flat.master.device


def test_flat_adamw_range_by_range_equals_one_pass():
    """The data-parallel trainer applies AdamW bucket by bucket as the gradient allreduces complete (begin_step / step_range /
    end_step): any cover of the flat buffer by ranges -- cut at arbitrary 64-element boundaries, also inside a parameter and
    across frozen ones -- must give bit-identical state to the one-pass step()."""
    from b200sd.train import FlatParams
    from b200sd.trainer import FlatAdamW
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import TINY_OVERRIDES
    _setup()
    res = []
    for mode in ("one_pass", "ranges"):
        torch.manual_seed(0)
        m = UNet2DConditionModel(**TINY_OVERRIDES).to(DEV)
        m.mid_block.resnets[0].conv1.weight.requires_grad_(False)      # a frozen region in the middle
        flat = FlatParams(m, torch.device(DEV))
        flat.attach_grads()
        opt = FlatAdamW(flat, lr=1e-3, weight_decay=0.1)
        g = torch.Generator(device=DEV).manual_seed(3)
        for step in range(2):
            flat.grad.copy_(torch.randn(flat.grad.shape, generator=g, device=DEV))
            if mode == "one_pass":
                opt.step(grad_scale=0.25, zero_grad=True)
            else:
                n = flat.grad.numel()
                cuts = sorted({0, n} | {int(c) // 64 * 64 for c in torch.linspace(0, n, 9)[1:-1].tolist()} | {64 * 1001, 64 * 1002})
                opt.begin_step()
                for a, b in reversed(list(zip(cuts[:-1], cuts[1:]))):      # back to front, as the backward completes them
                    opt.step_range(a, b, grad_scale=0.25, zero_grad=True)
                opt.end_step()
        res.append((flat.master.clone(), flat.wb.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), flat.grad.clone(), opt.steps))
    for a, b in zip(res[0][:5], res[1][:5]):
        assert torch.equal(a, b)
    assert res[0][5] == res[1][5] == 2


def test_trainer_steps_reduce_the_loss():
    """finetune_sd.py:453-494 loop body through b200sd.trainer.Trainer on one GPU: a fixed batch is (over)fitted."""
    from b200sd.schedulers import DDPMScheduler
    from b200sd.trainer import Trainer
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import TINY_OVERRIDES
    _setup()
    torch.manual_seed(0)
    unet = UNet2DConditionModel(**TINY_OVERRIDES).to(DEV)
    sched = DDPMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000)
    tr = Trainer(unet, sched, lr=2e-4, weight_decay=0.0)
    x, noise, ctx, t = _inputs(4, 32, 32, 64, 5)
    before = unet.conv_out.weight.detach().clone()
    losses = [float(tr.train_step(x, noise, t, ctx)) for _ in range(12)]
    assert losses[-1] < 0.7 * losses[0], losses
    assert not torch.equal(before, unet.conv_out.weight.detach())
    # gradient accumulation: two sync=False micro-steps + one sync step == one step on the 3x gradient
    g0 = unet.flat_gradients()
    assert float(g0.abs().max()) == 0.0
    tr.train_step(x, noise, t, ctx, sync=False)
    g1 = g0.clone()
    tr.train_step(x, noise, t, ctx, sync=False)
    assert torch.allclose(g0, 2 * g1, rtol=1e-3, atol=1e-4 * float(g1.abs().max()))   # fp32 atomics: run-to-run noise ~1e-6 of the max
    # the inference engine picks the updated weights up (eval forward == training forward on the same weights)
    unet.eval()
    with torch.no_grad():
        a = unet(x, t, ctx).sample
    unet.train()
    b = unet(x, t, ctx).sample.detach()
    rel = float((a - b).abs().max() / b.abs().max())
    assert rel <= 2e-2, rel


def test_trainer_optimizer_takes_the_reference_lr_scheduler_and_survives_a_rebuild_of_the_flat_buffers():
    """finetune_sd.py:421-422, 569-577: `scheduler.step()` after every optimizer step.  Trainer.optimizer() is the FlatAdamW (a
    torch.optim.Optimizer); unet.to() un-homes the parameters, the trainer rebuilds the flat buffers and re-binds the SAME optimizer
    object: moments, step count and the attached scheduler carry over."""
    from b200sd.schedulers import DDPMScheduler
    from b200sd.trainer import Trainer
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import TINY_OVERRIDES
    _setup()
    torch.manual_seed(0)
    unet = UNet2DConditionModel(**TINY_OVERRIDES).to(DEV)
    sched = DDPMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000)
    tr = Trainer(unet, sched, lr=2e-4, weight_decay=0.0)
    opt = tr.optimizer()
    assert isinstance(opt, torch.optim.Optimizer)
    lr_sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=8, eta_min=1e-6)
    x, noise, ctx, t = _inputs(2, 32, 32, 64, 5)
    losses, lrs = [], []
    for it in range(6):
        if it == 3:
            unet.to(DEV)                       # _apply: the parameters leave the flat master buffer
        losses.append(float(tr.train_step(x, noise, t, ctx)))
        lr_sched.step()
        lrs.append(opt.lr)
        assert tr.opt is opt and opt.steps == it + 1
    assert lrs == sorted(lrs, reverse=True) and lrs[-1] == lr_sched.get_last_lr()[0] < 1e-4
    assert float(opt.exp_avg.abs().max()) > 0 and losses[-1] < losses[0], losses
    sd = opt.state_dict()
    assert sd["state"][0]["step"] == 6 and sd["state"][0]["exp_avg"].numel() == opt.flat.master.numel()


def test_autograd_mode_accumulates_and_survives_save_load(tmp_path):
    """Default (autograd) gradient delivery: two backwards without zero_grad accumulate in param.grad like an eager
    module; the re-homed parameters still round-trip through save_pretrained / from_pretrained (finetune_sd.py:517-537)."""
    from b200sd import ops
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import TINY_OVERRIDES
    _setup()
    torch.manual_seed(0)
    unet = UNet2DConditionModel(**TINY_OVERRIDES).to(DEV).train()
    x, noise, ctx, t = _inputs(2, 32, 32, 64, 7)
    ops.mse_loss(unet(x, t, ctx).sample, noise).backward()
    g1 = {n: p.grad.clone() for n, p in unet.named_parameters()}
    assert all(p.grad.is_contiguous() or p.dim() == 4 for p in unet.parameters())
    ops.mse_loss(unet(x, t, ctx).sample, noise).backward()
    for n, p in unet.named_parameters():
        assert torch.allclose(p.grad, 2 * g1[n], rtol=1e-3, atol=1e-4 * float(g1[n].abs().max()) + 1e-12), n
    # a stock torch optimizer steps the re-homed parameters; the next forward sees the new weights
    opt = torch.optim.AdamW(unet.parameters(), lr=1e-3)
    with torch.no_grad():
        before = unet(x, t, ctx).sample.clone()
    opt.step()
    opt.zero_grad()
    with torch.no_grad():
        after = unet(x, t, ctx).sample
    assert float((after - before).abs().max()) > 0
    for safe in (False, True):
        d = tmp_path / ("st" if safe else "bin")
        unet.save_pretrained(str(d), safe_serialization=safe)
        again = UNet2DConditionModel.from_pretrained(str(d))
        for (n, p), (_, q) in zip(unet.state_dict().items(), again.state_dict().items()):
            assert torch.equal(p.detach().cpu(), q), n


def test_sd15_unet_gradients_batch8_and_frozen_context_gradient():
    """The benchmarked training shapes (BASELINE configs 3 and 4): full SD v1.5 UNet at batch 8 -- all 686 parameter gradients
    with the weights trainable, then the data-gradient-only plan (UNet frozen, finetune_sd.py:391-395) down to the context."""
    from b200sd import ops
    oracle, ours = _pair({})
    x, noise, ctx, t = _inputs(8, 64, 64, 768, 11)
    ref_loss, ref, ref_dctx = _ref_grads(oracle, x, noise, ctx, t, ctx_grad=True)
    oracle.zero_grad(set_to_none=True)
    loss = ops.mse_loss(ours(x, t, ctx).sample, noise)
    loss.backward()
    assert abs(float(loss) - ref_loss) <= 2e-2 * ref_loss
    worst, cos_all, bad = _compare(ref, {n: p.grad for n, p in ours.named_parameters()})
    print(f"sd15 batch-8 grads: worst per-tensor max-rel {worst:.4g}, global cosine {cos_all:.6f}, out-of-tolerance {len(bad)}")
    assert not bad, f"{len(bad)} tensors out of tolerance, e.g. {bad[:5]}"
    assert cos_all >= 0.9995, cos_all
    del ref
    torch.cuda.empty_cache()
    # config 4: frozen UNet, gradient into the (CLIP) context only
    ours.zero_grad(set_to_none=True)
    ours.requires_grad_(False).eval()
    c = ctx.clone().requires_grad_(True)
    ops.mse_loss(ours(x, t, c).sample, noise).backward()
    assert all(p.grad is None for p in ours.parameters())
    rel = float((c.grad - ref_dctx).abs().max() / ref_dctx.abs().max())
    cos = float(F.cosine_similarity(c.grad.flatten(), ref_dctx.flatten(), dim=0))
    print(f"sd15 batch-8 frozen-UNet context gradient: max-rel {rel:.4g}, cosine {cos:.6f}")
    assert rel <= 4e-2 and cos >= 0.999, (rel, cos)
