"""Scheduler facades with the diffusers 0.7.2 surface the reference touches:

    DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                  clip_sample=False, set_alpha_to_one=False)                 inference.py:386-387
    PNDMScheduler(..., skip_prk_steps=True)                                   utils.py:222-224
    DDPMScheduler.from_config(path, subfolder="scheduler").add_noise(...)     finetune_sd.py:335-336, 473

Host side: timestep tables and per-step scalar coefficients (fp64 from the fp32 abar table).
Device side: ONE fused kernel per step through the C ABI (b200sd_cfg_ddim_step /
b200sd_cfg_plms_step / b200sd_add_noise).  `step_cfg` additionally fuses the pipeline's
classifier-free-guidance combine into the same kernel.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import numpy as np
import torch

from . import ops


class SchedulerOutput(SimpleNamespace):
    pass


def _betas(num_train_timesteps, beta_start, beta_end, beta_schedule):
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    raise NotImplementedError(f"{beta_schedule} is not implemented")


class _SchedulerBase:
    _defaults: dict = {}

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", **kw):
        cfg = dict(self._defaults)
        cfg.update(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                   beta_schedule=beta_schedule)
        unknown = set(kw) - set(self._defaults)
        if unknown:
            raise TypeError(f"{type(self).__name__}: unexpected config keys {sorted(unknown)}")
        cfg.update(kw)
        self.config = SimpleNamespace(**cfg)
        self.num_train_timesteps = num_train_timesteps
        self.betas = _betas(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self._ac = self.alphas_cumprod.double().tolist()
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())
        self._tables = {}

    # -- diffusers config plumbing ---------------------------------------------------------
    @classmethod
    def from_config(cls, path_or_dict, subfolder=None, **overrides):
        if isinstance(path_or_dict, dict):
            cfg = dict(path_or_dict)
        else:
            p = path_or_dict if subfolder is None else os.path.join(path_or_dict, subfolder)
            with open(os.path.join(p, "scheduler_config.json")) as f:
                cfg = json.load(f)
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        allowed = {"num_train_timesteps", "beta_start", "beta_end", "beta_schedule", *cls._defaults}
        cfg = {k: v for k, v in cfg.items() if k in allowed}
        cfg.update(overrides)
        return cls(**cfg)

    from_pretrained = from_config

    def save_config(self, path):
        os.makedirs(path, exist_ok=True)
        d = dict(vars(self.config), _class_name=type(self).__name__)
        with open(os.path.join(path, "scheduler_config.json"), "w") as f:
            json.dump(d, f, indent=2)

    save_pretrained = save_config

    # -- shared behaviour ------------------------------------------------------------------
    def scale_model_input(self, sample, timestep=None):
        return sample

    def _noise_tables(self, device):
        key = str(device)
        if key not in self._tables:
            ac = self.alphas_cumprod.double()
            self._tables[key] = (ac.sqrt().float().to(device), (1 - ac).sqrt().float().to(device))
        return self._tables[key]

    def add_noise(self, original_samples, noise, timesteps):
        """sqrt(abar[t]) x0 + sqrt(1 - abar[t]) eps with a per-sample timestep (App. B.1)."""
        if original_samples.shape != noise.shape:
            raise ValueError("original_samples and noise must have the same shape")
        timesteps = timesteps.to(device=original_samples.device, dtype=torch.int64).reshape(-1)
        if timesteps.numel() == 1 and original_samples.shape[0] != 1:
            timesteps = timesteps.expand(original_samples.shape[0])
        if timesteps.numel() != original_samples.shape[0]:
            raise ValueError("timesteps must have one entry per sample")
        sa, sb = self._noise_tables(original_samples.device)
        return ops.add_noise(original_samples.contiguous(), noise.contiguous(), timesteps.contiguous(), sa, sb)

    def _alpha_prev(self, prev_t):
        return self._ac[prev_t] if prev_t >= 0 else self._final_alpha

    def _check_step_args(self, model_output, sample):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        if model_output.numel() != sample.numel():
            raise ValueError("model_output and sample must have the same number of elements")


class DDPMScheduler(_SchedulerBase):
    """Only the training-time surface the reference uses: add_noise / num_train_timesteps."""
    _defaults = dict(trained_betas=None, variance_type="fixed_small", clip_sample=True)


class DDIMScheduler(_SchedulerBase):
    _defaults = dict(trained_betas=None, clip_sample=True, set_alpha_to_one=True, steps_offset=0)

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._final_alpha = 1.0 if self.config.set_alpha_to_one else self._ac[0]
        self.final_alpha_cumprod = torch.tensor(self._final_alpha, dtype=torch.float32)
        if self.config.clip_sample:
            raise NotImplementedError("clip_sample=True is not on the reference path (inference.py:386-387)")

    def set_timesteps(self, num_inference_steps, device=None):
        if num_inference_steps > self.num_train_timesteps or num_inference_steps < 1:
            raise ValueError("num_inference_steps out of range")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts + self.config.steps_offset)

    def _coefs(self, timestep):
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t, a_p = self._ac[t], self._alpha_prev(prev_t)
        return a_t ** 0.5, (1 - a_t) ** 0.5, a_p ** 0.5, (1 - a_p) ** 0.5

    def step(self, model_output, timestep, sample, eta: float = 0.0, return_dict: bool = True, **kw):
        if eta != 0.0:
            raise NotImplementedError("eta != 0 is not on the reference path")
        self._check_step_args(model_output, sample)
        prev = ops.cfg_ddim_step(model_output.contiguous(), None, sample.contiguous(), 0.0, *self._coefs(timestep))
        return SchedulerOutput(prev_sample=prev) if return_dict else (prev,)

    def step_cfg(self, model_output_2b, timestep, sample, guidance_scale, out=None):
        """Fused `eps_u + s (eps_c - eps_u)` + DDIM update; model_output_2b is the (2B, ...) UNet output."""
        self._check_step_args(model_output_2b[: sample.shape[0]], sample)
        B = sample.shape[0]
        prev = ops.cfg_ddim_step(model_output_2b[:B], model_output_2b[B:], sample.contiguous(), guidance_scale,
                                 *self._coefs(timestep), out=out)
        return SchedulerOutput(prev_sample=prev)


class PNDMScheduler(_SchedulerBase):
    _defaults = dict(trained_betas=None, skip_prk_steps=False, set_alpha_to_one=False, steps_offset=0)

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        if not self.config.skip_prk_steps:
            raise NotImplementedError("only PLMS (skip_prk_steps=True) is on the reference path (utils.py:222-224)")
        self._final_alpha = 1.0 if self.config.set_alpha_to_one else self._ac[0]
        self.final_alpha_cumprod = torch.tensor(self._final_alpha, dtype=torch.float32)
        self.pndm_order = 4
        self.ets, self.counter, self.cur_sample = [], 0, None

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        _t = (np.arange(0, num_inference_steps) * ratio).round() + self.config.steps_offset
        plms = np.concatenate([_t[:-1], _t[-2:-1], _t[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets, self.counter, self.cur_sample = [], 0, None

    def _plms(self, eps_u, eps_c, timestep, sample, guidance, out=None):
        self._check_step_args(eps_u, sample)
        t = int(timestep)
        ratio = self.num_train_timesteps // self.num_inference_steps
        prev_t = t - ratio
        keep_eps = self.counter != 1
        if keep_eps:
            self.ets = self.ets[-3:]
            n_after = len(self.ets) + 1
        else:
            prev_t, t = t, t + ratio
            n_after = len(self.ets)
        if n_after == 1 and self.counter == 0:
            w, hist, x = [1.0], [], sample
            self.cur_sample = sample.clone()    # the caller may recycle its latent buffers (step_cfg(out=...))
        elif n_after == 1 and self.counter == 1:
            w, hist, x = [0.5, 0.5], [self.ets[-1]], self.cur_sample
            self.cur_sample = None
        elif n_after == 2:
            w, hist, x = [1.5, -0.5], [self.ets[-1]], sample
        elif n_after == 3:
            w, hist, x = [23 / 12, -16 / 12, 5 / 12], [self.ets[-1], self.ets[-2]], sample
        else:
            w, hist, x = [55 / 24, -59 / 24, 37 / 24, -9 / 24], [self.ets[-1], self.ets[-2], self.ets[-3]], sample
        a_t, a_p = self._ac[t], self._alpha_prev(prev_t)
        b_t, b_p = 1 - a_t, 1 - a_p
        cx = (a_p / a_t) ** 0.5
        ce = (a_p - a_t) / (a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5)
        eps_out = torch.empty_like(eps_u) if keep_eps else None
        if eps_u.dtype == torch.float16:      # fp16 pipelines: the eps history is kept in fp32
            eps_u, eps_c = eps_u.float(), (None if eps_c is None else eps_c.float())
            eps_out = torch.empty_like(eps_u) if keep_eps else None
            prev = ops.cfg_plms_step(eps_u, eps_c, x.float().contiguous(), hist, w, guidance, cx, ce, eps_out=eps_out).to(x.dtype)
            if out is not None:
                prev = out.copy_(prev)
        else:
            prev = ops.cfg_plms_step(eps_u, eps_c, x.contiguous(), hist, w, guidance, cx, ce, out=out, eps_out=eps_out)
        if keep_eps:
            self.ets.append(eps_out)
        self.counter += 1
        return SchedulerOutput(prev_sample=prev)

    def plms_plan(self):
        """The whole call sequence of one sampling run as data: one row per UNet call for b200sd_cfg_plms_step_table
        (w0 w1 w2 w3 | cx ce | ring slots of ets[-1], ets[-2], ets[-3] | slot this call's eps goes to | x_from_saved | save_x).
        It is `_plms`'s counter logic run once on the host with a 4-slot eps ring instead of tensors, so that
        sampler.CapturedSampler can replay every call as the same CUDA graph (the step index lives on the device)."""
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps first")
        ratio = self.num_train_timesteps // self.num_inference_steps
        rows, ets = [], []
        for counter, timestep in enumerate(self.timesteps.tolist()):
            t = int(timestep)
            prev_t = t - ratio
            keep_eps = counter != 1
            if keep_eps:
                ets = ets[-3:]
                n_after = len(ets) + 1
            else:
                prev_t, t = t, t + ratio
                n_after = len(ets)
            from_saved = save_x = 0.0
            if n_after == 1 and counter == 0:
                w, hist, save_x = [1.0], [], 1.0
            elif n_after == 1 and counter == 1:
                w, hist, from_saved = [0.5, 0.5], [ets[-1]], 1.0
            elif n_after == 2:
                w, hist = [1.5, -0.5], [ets[-1]]
            elif n_after == 3:
                w, hist = [23 / 12, -16 / 12, 5 / 12], [ets[-1], ets[-2]]
            else:
                w, hist = [55 / 24, -59 / 24, 37 / 24, -9 / 24], [ets[-1], ets[-2], ets[-3]]
            a_t, a_p = self._ac[t], self._alpha_prev(prev_t)
            b_t, b_p = 1 - a_t, 1 - a_p
            cx = (a_p / a_t) ** 0.5
            ce = (a_p - a_t) / (a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5)
            slot = -1
            if keep_eps:
                slot = next(k for k in range(4) if k not in ets)       # ets holds <= 3 live slots here
                ets.append(slot)
            rows.append(w + [0.0] * (4 - len(w)) + [cx, ce] + [float(h) for h in hist] + [-1.0] * (3 - len(hist))
                        + [float(slot), from_saved, save_x])
        return rows

    def step(self, model_output, timestep, sample, return_dict: bool = True, **kw):
        out = self._plms(model_output.contiguous(), None, timestep, sample, 0.0)
        return out if return_dict else (out.prev_sample,)

    def step_cfg(self, model_output_2b, timestep, sample, guidance_scale, out=None):
        B = sample.shape[0]
        return self._plms(model_output_2b[:B], model_output_2b[B:], timestep, sample, guidance_scale, out=out)
