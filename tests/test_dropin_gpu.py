"""Drop-in boundary on the GPU: the ways the reference's scripts actually hold the facade classes.

    accelerator.prepare(unet)  == DistributedDataParallel wrap          finetune_sd.py:363, 386
    with torch.autocast('cuda'):  unet(...)                              finetune_sd.py:453
    unet.to(device, dtype=torch.float16) on the frozen copy              finetune_sd.py:393
    StableDiffusionPipeline.from_pretrained(torch_dtype=torch.float16)   inference.py:406, 425; utils.py:189, 249
    sampling between training steps                                      finetune_sd.py:264-271

Each case is checked against the fp32 oracle (or against our own fp32-input result where the oracle has no fp16 path)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair(train=False):
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import TINY_OVERRIDES, make_oracle_unet
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    oracle = make_oracle_unet(seed=0, **TINY_OVERRIDES)
    ours = UNet2DConditionModel(**TINY_OVERRIDES)
    ours.load_state_dict(oracle.state_dict(), strict=True)
    ours = ours.to(DEV)
    return oracle, (ours.train() if train else ours.eval())


def _inputs(N, seed=3, hw=32, ctx_dim=64):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(N, 4, hw, hw, generator=g), torch.randn(N, 4, hw, hw, generator=g),
            torch.randn(N, 77, ctx_dim, generator=g), torch.randint(0, 1000, (N,), generator=g))


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max())


def test_fp16_pipeline_dtypes_are_accepted():
    """An fp16 pipeline (inference.py:406): fp16 UNet weights, fp16 latents / context in, fp16 noise prediction out, and the
    schedulers step fp16 tensors -- results equal the fp32-input path up to fp16 rounding of the inputs / outputs."""
    from b200sd.schedulers import DDIMScheduler, DDPMScheduler, PNDMScheduler
    from b200sd import ops
    oracle, unet = _pair()
    x, noise, ctx, t = _inputs(2)
    with torch.no_grad():
        want = oracle(x, 500, ctx).sample
        unet16 = unet.to(DEV, dtype=torch.float16)                       # finetune_sd.py:393
        assert unet16.dtype == torch.float16
        got = unet16(x.to(DEV).half(), 500, ctx.to(DEV).half()).sample
    assert got.dtype == torch.float16
    assert _rel(got.cpu(), want) <= 3e-2
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    eps32, lat32 = noise.to(DEV), x.to(DEV)
    for sch in (DDIMScheduler(clip_sample=False, set_alpha_to_one=False, **kw), PNDMScheduler(skip_prk_steps=True, **kw)):
        sch.set_timesteps(50)
        a = sch.step(eps32, int(sch.timesteps[0]), lat32).prev_sample
        sch.set_timesteps(50)
        b = sch.step(eps32.half(), int(sch.timesteps[0]), lat32.half()).prev_sample
        assert b.dtype == torch.float16 and _rel(b, a) <= 2e-3, type(sch).__name__
        sch.set_timesteps(50)
        c = sch.step_cfg(torch.cat([eps32, eps32 * 0.5]).half(), int(sch.timesteps[0]), lat32.half(), 7.5).prev_sample
        sch.set_timesteps(50)
        d = sch.step_cfg(torch.cat([eps32, eps32 * 0.5]), int(sch.timesteps[0]), lat32, 7.5).prev_sample
        assert c.dtype == torch.float16 and _rel(c, d) <= 4e-3, type(sch).__name__
    ddpm = DDPMScheduler(num_train_timesteps=1000, **kw)
    n32 = ddpm.add_noise(lat32, eps32, t.to(DEV))
    n16 = ddpm.add_noise(lat32.half(), eps32.half(), t.to(DEV))
    assert n16.dtype == torch.float16 and _rel(n16, n32) <= 2e-3
    l32 = ops.mse_loss(eps32, lat32)
    p16 = eps32.half().requires_grad_(True)
    l16 = ops.mse_loss(p16, lat32.half())
    l16.backward()
    assert abs(float(l16) - float(l32)) <= 2e-3 * float(l32)
    assert p16.grad.dtype == torch.float16 and _rel(p16.grad, 2 * (eps32 - lat32) / eps32.numel()) <= 5e-3


def test_frozen_fp16_unet_gives_the_context_gradient():
    """finetune_sd.py:391-395 + 477-494 (train_text_encoder): UNet frozen and cast to fp16, the gradient flows through it into
    the text context."""
    from b200sd import ops
    oracle, unet = _pair()
    x, noise, ctx, t = _inputs(2)
    c_ref = ctx.clone().requires_grad_(True)
    F.mse_loss(oracle(x, t, c_ref).sample, noise).backward()
    unet = unet.requires_grad_(False).to(DEV, dtype=torch.float16)
    c = ctx.to(DEV).half().requires_grad_(True)
    ops.mse_loss(unet(x.to(DEV).half(), t.to(DEV), c).sample, noise.to(DEV).half()).backward()
    assert c.grad is not None and c.grad.dtype == torch.float16
    cos = float(F.cosine_similarity(c.grad.float().flatten().cpu(), c_ref.grad.flatten(), dim=0))
    assert cos >= 0.995, cos
    # a second step reuses the flat state instead of rebuilding it
    flat = unet._flat
    c.grad = None
    ops.mse_loss(unet(x.to(DEV).half(), t.to(DEV), c).sample, noise.to(DEV).half()).backward()
    assert unet._flat is flat


def test_ddp_wrap_and_autocast_like_accelerate():
    """accelerator.prepare(unet) -> DistributedDataParallel (finetune_sd.py:363, 386), forward under torch.autocast('cuda')
    with fp16 latents and context (finetune_sd.py:453): gradients arrive in param.grad through DDP's hooks and match the oracle."""
    import torch.distributed as dist
    oracle, unet = _pair(train=True)
    x, noise, ctx, t = _inputs(2)
    oracle.zero_grad(set_to_none=True)
    F.mse_loss(oracle(x, t, ctx).sample, noise).backward()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
    try:
        ddp = torch.nn.parallel.DistributedDataParallel(unet, device_ids=[0])
        with torch.autocast("cuda", dtype=torch.float16):
            pred = ddp(x.to(DEV).half(), t.to(DEV), ctx.to(DEV).half()).sample
            loss = F.mse_loss(pred.float(), noise.to(DEV), reduction="none").mean([1, 2, 3]).mean()     # finetune_sd.py:483-484
        loss.backward()
        ref = torch.cat([p.grad.flatten() for _, p in oracle.named_parameters()])
        named = dict(unet.named_parameters())
        got = torch.cat([named[n].grad.flatten().cpu() for n, _ in oracle.named_parameters()])
        cos = float(F.cosine_similarity(got.double(), ref.double(), dim=0))
        assert cos >= 0.995, cos
    finally:
        if created:
            dist.destroy_process_group()


def test_sampling_between_training_steps_sees_the_new_weights():
    """finetune_sd.py:264-271 samples with the model being trained: eval forward -> train_step (fused flat AdamW: no tensor
    version bump) -> eval forward must use the UPDATED weights and agree with a fresh model loaded from state_dict()."""
    from b200sd.schedulers import DDPMScheduler
    from b200sd.trainer import Trainer
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import TINY_OVERRIDES
    _, unet = _pair()
    x, noise, ctx, t = (v.to(DEV) for v in _inputs(2))
    with torch.no_grad():
        before = unet(x, t, ctx).sample.clone()
    sched = DDPMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000)
    tr = Trainer(unet, sched, lr=1e-3, weight_decay=0.0)
    for _ in range(3):
        tr.train_step(x, noise, t, ctx)
    unet.eval()
    with torch.no_grad():
        after = unet(x, t, ctx).sample.clone()
    assert float((after - before).abs().max()) > 1e-3 * float(before.abs().max())
    fresh = UNet2DConditionModel(**TINY_OVERRIDES)
    fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in unet.state_dict().items()}, strict=True)
    fresh = fresh.to(DEV).eval()
    with torch.no_grad():
        want = fresh(x, t, ctx).sample
    assert _rel(after, want) <= 1e-3


def test_two_forwards_before_backward_is_an_error_not_a_wrong_gradient():
    from b200sd import ops
    _, unet = _pair(train=True)
    x, noise, ctx, t = (v.to(DEV) for v in _inputs(2))
    l1 = ops.mse_loss(unet(x, t, ctx).sample, noise)
    l2 = ops.mse_loss(unet(x, t, ctx).sample, noise)
    l2.backward()
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        l1.backward()


def test_accumulation_averages_and_frozen_parameters_do_not_decay():
    """accelerate semantics (finetune_sd.py:454-458, 494): k accumulated micro-steps are averaged, so 2 x sync=False + 1 x sync on
    one batch == one plain step on that batch; parameters with requires_grad=False get neither update nor weight decay."""
    from b200sd.schedulers import DDPMScheduler
    from b200sd.trainer import Trainer
    sched = DDPMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000)
    x, noise, ctx, t = (v.to(DEV) for v in _inputs(2))
    deltas = []
    for micro in (1, 3):
        _, unet = _pair(train=True)
        unet.conv_in.weight.requires_grad_(False)
        frozen_before = unet.conv_in.weight.detach().clone()
        before = unet.conv_out.weight.detach().clone()
        # eps = 1 makes AdamW's first step ~ lr * g, i.e. sensitive to the SCALE of the gradient (with the default eps the
        # first Adam step is lr * sign(g) whatever the scale)
        tr = Trainer(unet, sched, lr=1e-1, eps=1.0, weight_decay=0.1)
        for i in range(micro):
            tr.train_step(x, noise, t, ctx, sync=(i == micro - 1))
        assert torch.equal(unet.conv_in.weight.detach(), frozen_before)
        deltas.append(unet.conv_out.weight.detach() - before)
    assert float(deltas[0].abs().max()) > 0
    assert _rel(deltas[1], deltas[0]) <= 1e-2


def test_pipeline_call_matches_the_oracle_loop():
    """inference.py:342-351: pipeline(..., height, width, num_inference_steps, guidance_scale, latents=fixed) -- through the
    StableDiffusionPipeline wrapper with precomputed text embeddings (CLIP is a neighbour) vs the oracle's App. B.4 loop."""
    from b200sd import StableDiffusionPipeline
    from b200sd.schedulers import DDIMScheduler
    from oracle import schedulers_ref as R
    oracle, unet = _pair()
    g = torch.Generator().manual_seed(42)
    lat = torch.randn(2, 4, 32, 32, generator=g)
    ctx2 = torch.randn(4, 77, 64, generator=g)
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False)
    pipe = StableDiffusionPipeline(unet=unet, scheduler=DDIMScheduler(**kw), safety_checker=None).to(DEV)
    out = pipe(prompt_embeds=ctx2.to(DEV), height=256, width=256, num_inference_steps=6, guidance_scale=7.5, latents=lat.to(DEV),
               output_type="latent")
    with torch.no_grad():
        want = R.denoise_loop(oracle, R.DDIMSchedulerRef(clip_sample=False, set_alpha_to_one=False), lat, ctx2, 6, 7.5)
    cos = float(F.cosine_similarity(out.images.float().cpu().flatten(), want.flatten(), dim=0))
    assert out.images.shape == (2, 4, 32, 32) and cos >= 0.999, cos
    with pytest.raises(ValueError, match="Unexpected latents shape"):
        pipe(prompt_embeds=ctx2.to(DEV), height=256, width=256, latents=lat[:, :, :16].to(DEV))


def test_reference_train_unet_loop_body_with_all_four_models():
    """The loop body of finetune_sd.py:453-494 in train_unet mode, every model b200sd's own (reduced widths):
        latents = vae.encode(batch["pixel_values"]).latent_dist.sample() * 0.18215        :460-462  (frozen fp16 VAE)
        noise, timesteps; noisy_latents = noise_scheduler.add_noise(...)                  :465-474
        encoder_hidden_states = text_encoder(batch["input_ids"])[0]                       :477      (frozen fp16 CLIP)
        noise_pred = unet(noisy_latents, timesteps, encoder_hidden_states).sample         :480-481
        loss = F.mse_loss(...).mean([1,2,3]).mean(); backward; optimizer.step()           :483-494, 569-570
    The loss must equal the chain of oracles on the same posterior noise, and a few steps must reduce it."""
    from b200sd.clip import CLIPTextModel
    from b200sd.schedulers import DDPMScheduler
    from b200sd.vae import AutoencoderKL
    from oracle import schedulers_ref as R
    from oracle.clip_ref import make_oracle_clip
    from oracle.vae_ref import TINY_VAE_OVERRIDES, make_oracle_vae
    o_unet, unet = _pair(train=True)
    clip_kw = dict(vocab_size=1000, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=1)
    o_clip, o_vae = make_oracle_clip(seed=0, **clip_kw), make_oracle_vae(seed=0, **TINY_VAE_OVERRIDES)
    clip = CLIPTextModel(**clip_kw)
    clip.load_state_dict(o_clip.state_dict(), strict=True)
    clip = clip.to(DEV, dtype=torch.float16).requires_grad_(False).eval()
    vae = AutoencoderKL(**TINY_VAE_OVERRIDES)
    vae.load_state_dict(o_vae.state_dict(), strict=True)
    vae = vae.to(DEV, dtype=torch.float16).requires_grad_(False).eval()
    g = torch.Generator().manual_seed(9)
    B = 2
    pixels = torch.randn(B, 3, 256, 256, generator=g).clamp(-1, 1)
    ids = torch.randint(2, 990, (B, 77), generator=g)
    t = torch.tensor([150, 720])
    sched = DDPMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000)
    opt = torch.optim.AdamW(unet.parameters(), lr=2e-4)
    losses = []
    for it in range(4):
        with torch.autocast("cuda", dtype=torch.float16):
            post = vae.encode(pixels.to(DEV).half()).latent_dist
            gen = torch.Generator(device=DEV).manual_seed(100)
            latents = post.sample(generator=gen) * 0.18215
            noise = torch.randn(latents.shape, generator=torch.Generator(device=DEV).manual_seed(200), device=DEV).to(latents.dtype)
            noisy = sched.add_noise(latents, noise, t.to(DEV))
            ehs = clip(ids.to(DEV))[0]
            pred = unet(noisy, t.to(DEV), ehs).sample
            loss = F.mse_loss(pred.float(), noise.float(), reduction="none").mean([1, 2, 3]).mean()
        if it == 0:
            with torch.no_grad():
                eps = torch.randn(latents.shape, generator=torch.Generator(device=DEV).manual_seed(100), device=DEV).cpu()
                lat_ref = o_vae.encode(pixels).latent_dist.sample(noise=eps) * 0.18215
                noisy_ref = R.DDPMSchedulerRef().add_noise(lat_ref, noise.float().cpu(), t)
                want = F.mse_loss(o_unet(noisy_ref, t, o_clip(ids)[0]).sample, noise.float().cpu())
            assert abs(float(loss) - float(want)) <= 3e-2 * float(want), (float(loss), float(want))
        loss.backward()
        opt.step()
        opt.zero_grad()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses
