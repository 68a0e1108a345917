"""CPU tests of the oracle (the checker): structural invariants of the public SD v1.x checkpoint, the
closed-form known answers of SURVEY.md App. B.5, committed golden vectors, and cross-implementation
checks.  (The pins against diffusers 0.7.2's own known-answer vectors are in tests/test_oracle_diffusers_kat.py.)"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import schedulers_ref as R
from oracle import unet_ref as U

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sd15_kat.json")))


@pytest.fixture(scope="module")
def tiny():
    return U.make_oracle_unet(seed=0, **U.TINY_OVERRIDES)


def test_sd15_parameter_and_tensor_counts():
    with torch.device("meta"):
        m = U.UNet2DConditionModelRef()
    assert sum(p.numel() for p in m.parameters()) == 859_520_964
    sd = m.state_dict()
    assert len(sd) == 686
    for k in ("conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.0.resnets.0.time_emb_proj.bias",
              "down_blocks.1.resnets.0.conv_shortcut.weight", "down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k.weight",
              "down_blocks.2.downsamplers.0.conv.weight", "mid_block.attentions.0.proj_in.weight",
              "up_blocks.3.attentions.2.transformer_blocks.0.ff.net.0.proj.bias", "up_blocks.0.upsamplers.0.conv.bias",
              "up_blocks.1.attentions.0.transformer_blocks.0.ff.net.2.weight", "conv_norm_out.weight", "conv_out.bias"):
        assert k in sd, k
    assert sd["mid_block.attentions.0.proj_in.weight"].shape == (1280, 1280, 1, 1)
    assert sd["up_blocks.1.resnets.2.conv1.weight"].shape == (1280, 1920, 3, 3)
    assert "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.bias" not in sd


def test_scheduler_known_answers():
    s = R.DDIMSchedulerRef(clip_sample=False, set_alpha_to_one=False)
    kat = {0: 0.99914998, 1: 0.99829602, 20: 0.98131430, 500: 0.27633247, 980: 0.00584378, 981: 0.00577550, 999: 0.00466010}
    for i, v in kat.items():
        assert abs(float(s.alphas_cumprod[i]) - v) < 2e-7 * max(1, v / 1e-3), (i, float(s.alphas_cumprod[i]))
        assert float(s.alphas_cumprod[i]) == GOLD["alphas_cumprod"][str(i)]
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(980, -1, -20)) == GOLD["ddim_timesteps_50"]
    p = R.PNDMSchedulerRef(skip_prk_steps=True, steps_offset=1)
    p.set_timesteps(50)
    ts = p.timesteps.tolist()
    assert len(ts) == 51 and ts[:4] == [981, 961, 961, 941] and ts[-2:] == [21, 1] and ts == GOLD["plms_timesteps_50_offset1"]


def test_timestep_embedding_known_answer():
    e = U.timestep_embedding(torch.tensor([980]))
    torch.testing.assert_close(e[0, :3], torch.tensor([0.98439258, 0.01940880, 0.99800313]), rtol=0, atol=2e-6)
    torch.testing.assert_close(e[0, 160:163], torch.tensor([-0.17598660, 0.99981165, 0.06316458]), rtol=0, atol=2e-6)
    assert e[0, :3].tolist() == GOLD["temb_980_cos_0_3"]


def test_golden_scheduler_vectors():
    s = R.DDIMSchedulerRef(clip_sample=False, set_alpha_to_one=False)
    s.set_timesteps(50)
    g = GOLD["ddim_step_t500"]
    x, eps = torch.tensor(g["x"]).view(1, 4, 4, 4), torch.tensor(g["eps"]).view(1, 4, 4, 4)
    torch.testing.assert_close(s.step(eps, 500, x).prev_sample.flatten(), torch.tensor(g["prev"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(R.DDPMSchedulerRef().add_noise(x, eps, torch.tensor([333])).flatten(),
                               torch.tensor(GOLD["add_noise_t333"]), rtol=1e-6, atol=1e-7)
    p = R.PNDMSchedulerRef(skip_prk_steps=True, steps_offset=1)
    p.set_timesteps(50)
    cur = x.clone()
    for i, tt in enumerate(p.timesteps[:6]):
        ee = torch.randn(1, 4, 4, 4, generator=torch.Generator().manual_seed(50 + i))
        cur = p.step(ee, tt, cur).prev_sample
        torch.testing.assert_close(cur.flatten(), torch.tensor(GOLD["plms_first6"][i]), rtol=1e-5, atol=1e-6)


def test_ddim_step_closed_form():
    """x' = sqrt(a_p/a_t) x + (sqrt(1-a_p) - sqrt(a_p (1-a_t)/a_t)) eps."""
    s = R.DDIMSchedulerRef(clip_sample=False, set_alpha_to_one=False)
    s.set_timesteps(50)
    x, eps = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    for t in (980, 500, 0):
        a_t = s.alphas_cumprod[t].double()
        a_p = (s.alphas_cumprod[t - 20] if t >= 20 else s.alphas_cumprod[0]).double()
        want = (a_p / a_t).sqrt() * x.double() + ((1 - a_p).sqrt() - (a_p * (1 - a_t) / a_t).sqrt()) * eps.double()
        torch.testing.assert_close(s.step(eps, t, x).prev_sample.double(), want, rtol=1e-5, atol=1e-5)


def test_attention_matches_sdpa():
    torch.manual_seed(0)
    a = U.CrossAttention(64, 48, 4)
    x, c = torch.randn(2, 50, 64), torch.randn(2, 7, 48)
    q = a.to_q(x).view(2, 50, 4, 16).transpose(1, 2)
    k = a.to_k(c).view(2, 7, 4, 16).transpose(1, 2)
    v = a.to_v(c).view(2, 7, 4, 16).transpose(1, 2)
    want = a.to_out[0](F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(2, 50, 64))
    torch.testing.assert_close(a(x, c), want, rtol=1e-4, atol=1e-5)


def test_conv_matches_unfold_matmul():
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(6, 10, 3, padding=1)
    x = torch.randn(2, 6, 9, 7)
    cols = F.unfold(x, 3, padding=1)
    want = (conv.weight.view(10, -1) @ cols + conv.bias[None, :, None]).view(2, 10, 9, 7)
    torch.testing.assert_close(conv(x), want, rtol=1e-4, atol=1e-5)


def test_tiny_unet_forward_shapes_and_timestep_forms(tiny):
    x, ctx = torch.randn(2, 4, 16, 16), torch.randn(2, 77, 64)
    with torch.no_grad():
        a = tiny(x, 10, ctx).sample
        b = tiny(x, torch.tensor(10), ctx).sample
        c = tiny(x, torch.tensor([10, 10]), ctx, return_dict=False)[0]
    assert a.shape == (2, 4, 16, 16)
    torch.testing.assert_close(a, b)
    torch.testing.assert_close(a, c)


def test_cfg_and_loop_and_mse(tiny):
    eps2 = torch.randn(4, 4, 8, 8)
    torch.testing.assert_close(R.cfg_combine(eps2, 7.5), eps2[:2] + 7.5 * (eps2[2:] - eps2[:2]))
    lat, ctx2 = torch.randn(1, 4, 16, 16), torch.randn(2, 77, 64)
    with torch.no_grad():
        out = R.denoise_loop(tiny, R.DDIMSchedulerRef(clip_sample=False, set_alpha_to_one=False), lat, ctx2, 3, 7.5)
    assert out.shape == lat.shape and torch.isfinite(out).all()
    p, t = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8)
    torch.testing.assert_close(R.mse_loss_ref(p, t), ((p - t) ** 2).mean())
