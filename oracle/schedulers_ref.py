"""ORACLE (test infrastructure, not product code) -- diffusers 0.7.2 scheduler / pipeline-loop math
restated in plain PyTorch + Python floats.

PARITY PINNED against the known answers diffusers 0.7.2 holds in its own test suite
(tests/test_scheduler.py: the DDIM and PNDM `test_full_loop_*` sums / means, transcribed into
tests/golden/diffusers_0_7_2_kat.json; checked by tests/test_oracle_diffusers_kat.py).  `diffusers==0.7.2`
(env.yaml:112) itself is un-vendored and not installable here and the reference holds no scheduler tests;
`add_noise` has no diffusers known answer and stays pinned by its closed form only.  Anchors are the
reference call sites:

  * DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
    clip_sample=False, set_alpha_to_one=False)                     -- inference.py:386-387
  * PNDMScheduler(..., skip_prk_steps=True)                         -- utils.py:222-224
  * DDPMScheduler.from_config(...).add_noise(latents, noise, t)     -- finetune_sd.py:335-336, 473-474
  * pipeline(..., num_inference_steps=50, guidance_scale=7.5)       -- inference.py:175-176
  * F.mse_loss(pred, noise, "none").mean([1,2,3]).mean()            -- finetune_sd.py:483-484

Known answers are checked in tests/test_oracle_schedulers.py.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


def make_betas(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear"):
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    raise NotImplementedError(beta_schedule)


class _Base:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 beta_schedule="scaled_linear", **kw):
        self.num_train_timesteps = num_train_timesteps
        self.betas = make_betas(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, **kw)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        """App. B.1: sqrt(abar[t]) x0 + sqrt(1-abar[t]) eps, table cast to x0's dtype first."""
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original_samples + sb * noise


class DDPMSchedulerRef(_Base):
    pass


class DDIMSchedulerRef(_Base):
    def __init__(self, clip_sample=True, set_alpha_to_one=True, steps_offset=0, **kw):
        super().__init__(clip_sample=clip_sample, set_alpha_to_one=set_alpha_to_one,
                         steps_offset=steps_offset, **kw)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, self.num_train_timesteps)[::-1].copy())

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts + self.config.steps_offset)

    def step(self, model_output, timestep, sample, eta: float = 0.0):
        """App. B.2 (eta = 0)."""
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            x0 = torch.clamp(x0, -1, 1)
        assert eta == 0.0, "oracle restates the reference's eta=0 path only"
        direction = (1 - a_p) ** 0.5 * model_output
        prev = a_p ** 0.5 * x0 + direction
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)


class PNDMSchedulerRef(_Base):
    def __init__(self, skip_prk_steps=False, set_alpha_to_one=False, steps_offset=0, **kw):
        super().__init__(skip_prk_steps=skip_prk_steps, set_alpha_to_one=set_alpha_to_one,
                         steps_offset=steps_offset, **kw)
        assert skip_prk_steps, "oracle restates PLMS (skip_prk_steps=True) only (utils.py:222-224)"
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.pndm_order = 4
        self.ets = []
        self.counter = 0
        self.cur_sample = None
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, self.num_train_timesteps)[::-1].copy())

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        _t = (np.arange(0, num_inference_steps) * ratio).round() + self.config.steps_offset
        plms = np.concatenate([_t[:-1], _t[-2:-1], _t[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def step(self, model_output, timestep, sample):
        """App. B.3 step_plms."""
        t = int(timestep)
        ratio = self.num_train_timesteps // self.num_inference_steps
        prev_t = t - ratio
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_t = t
            t = t + ratio
        if len(self.ets) == 1 and self.counter == 0:
            e = model_output
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            e = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            e = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            e = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            e = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        prev = self._get_prev_sample(sample, t, prev_t, e)
        self.counter += 1
        return SimpleNamespace(prev_sample=prev)

    def _get_prev_sample(self, sample, t, prev_t, e):
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        b_p = 1 - a_p
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return sample_coeff * sample - (a_p - a_t) * e / denom


def pndm_prk_timesteps(sch: "PNDMSchedulerRef"):
    """diffusers 0.7.2 scheduling_pndm.py `set_timesteps`, skip_prk_steps=False branch: (prk_timesteps, plms_timesteps)"""
    n = sch.num_inference_steps
    ratio = sch.num_train_timesteps // n
    _t = (np.arange(0, n) * ratio).round() + sch.config.steps_offset
    prk = np.array(_t[-sch.pndm_order:]).repeat(2) + np.tile(np.array([0, ratio // 2]), sch.pndm_order)
    prk = (prk[:-1].repeat(2)[1:-1])[::-1].copy().astype(np.int64)
    plms = _t[:-3][::-1].copy().astype(np.int64)
    return prk, plms


def pndm_prk_warmup(sch: "PNDMSchedulerRef", model, sample):
    """diffusers 0.7.2 scheduling_pndm.py `set_timesteps` (skip_prk_steps=False branch) + `step_prk`: the 12
    Runge-Kutta calls that precede PLMS when PRK is not skipped.  NOT on the reference path (utils.py:222-224 sets
    skip_prk_steps=True); restated only because diffusers' own known-answer test of `step_plms` /
    `_get_prev_sample` (tests/test_scheduler.py PNDMSchedulerTest.full_loop) enters PLMS through it.
    Leaves `sch` as diffusers would (3 saved eps, counter 12) and returns (sample, plms_timesteps)."""
    n = sch.num_inference_steps
    ratio = sch.num_train_timesteps // n
    prk, plms = pndm_prk_timesteps(sch)
    cur_out, cur_sample, ets = 0, None, []
    for c, t in enumerate(prk):
        t = int(t)
        out = model(sample, t)
        prev_t = t - (0 if c % 2 else ratio // 2)
        t0 = int(prk[c // 4 * 4])
        if c % 4 == 0:
            cur_out = cur_out + out / 6
            ets.append(out)
            cur_sample = sample
        elif c % 4 in (1, 2):
            cur_out = cur_out + out / 3
        else:
            out = cur_out + out / 6
            cur_out = 0
        sample = sch._get_prev_sample(cur_sample, t0, prev_t, out)
    sch.ets, sch.counter = ets, len(prk)
    return sample, plms


def cfg_combine(eps2, guidance_scale: float):
    """App. B.4: eps_u + s (eps_c - eps_u) after chunk(2)."""
    eps_u, eps_c = eps2.chunk(2)
    return eps_u + guidance_scale * (eps_c - eps_u)


def denoise_loop(unet, scheduler, latents, ctx2, num_inference_steps=50, guidance_scale=7.5, record=None):
    """App. B.4: the StableDiffusionPipeline.__call__ denoise loop (tokeniser/CLIP/VAE excluded).
    ctx2 = cat([uncond, cond]) with shape (2B, 77, 768)."""
    scheduler.set_timesteps(num_inference_steps)
    latents = latents * scheduler.init_noise_sigma
    for t in scheduler.timesteps:
        x2 = torch.cat([latents] * 2)
        x2 = scheduler.scale_model_input(x2, t)
        eps = unet(x2, t, ctx2).sample
        eps = cfg_combine(eps, guidance_scale)
        if record is not None:
            record.append(eps)
        latents = scheduler.step(eps, t, latents).prev_sample
    return latents


def mse_loss_ref(pred, target):
    """finetune_sd.py:483-484."""
    return torch.nn.functional.mse_loss(pred, target, reduction="none").mean([1, 2, 3]).mean()
