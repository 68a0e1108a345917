"""The denoising loop of StableDiffusionPipeline.__call__ -- the part of the
pipeline that is on the hot path.  Tokeniser / CLIP / VAE stay outside (neighbours, out of scope):
callers pass the (2B, 77, 768) context `cat([uncond, cond])` and the initial latents.

Reference call sites: inference.py:175-176, 342-351; finetune_sd.py:264-271."""
from __future__ import annotations

import torch


@torch.no_grad()
def denoise_loop(unet, scheduler, latents, encoder_hidden_states_2b, num_inference_steps=50, guidance_scale=7.5,
                 record=None):
    """latents (B,4,h,w) -> denoised latents (B,4,h,w).  One UNet call at batch 2B per step, then ONE
    fused kernel for `eps_u + s (eps_c - eps_u)` + the scheduler update."""
    if encoder_hidden_states_2b.shape[0] != 2 * latents.shape[0]:
        raise ValueError("encoder_hidden_states must be cat([uncond, cond]) with batch 2B")
    scheduler.set_timesteps(num_inference_steps)
    sampler = _captured_sampler(unet, scheduler, latents, encoder_hidden_states_2b, guidance_scale) if record is None else None
    if sampler is not None:
        return sampler.run(latents, encoder_hidden_states_2b).to(latents.dtype)
    latents = (latents * scheduler.init_noise_sigma).float().contiguous()
    x2 = torch.empty((2 * latents.shape[0],) + tuple(latents.shape[1:]), dtype=latents.dtype, device=latents.device)
    B = latents.shape[0]
    for t in scheduler.timesteps.tolist():
        x2[:B].copy_(latents)
        x2[B:].copy_(latents)
        eps2 = unet(scheduler.scale_model_input(x2, t), t, encoder_hidden_states_2b).sample
        if record is not None:
            record.append(eps2[:B] + guidance_scale * (eps2[B:] - eps2[:B]))
        latents = scheduler.step_cfg(eps2.float() if eps2.dtype not in (torch.float32, torch.bfloat16) else eps2, t,
                                     latents, guidance_scale).prev_sample
    return latents


def _captured_sampler(unet, scheduler, latents, ctx2, guidance_scale):
    """The whole-step CUDA graph (sampler.CapturedSampler) when the combination is covered: b200sd UNet on its bf16 plan with
    CUDA graphs on, DDIM or PLMS, CUDA tensors.  Cached on the UNet per (geometry, schedule, guidance); None -> the per-call loop."""
    from .schedulers import DDIMScheduler, PNDMScheduler
    from .unet import UNet2DConditionModel
    if not (isinstance(unet, UNet2DConditionModel) and isinstance(scheduler, (DDIMScheduler, PNDMScheduler)) and latents.is_cuda
            and unet._precision == "bf16" and unet.use_cuda_graph and not unet.training):
        return None
    from .sampler import CapturedSampler, default_lanes
    B, _, h, w = latents.shape
    key = (type(scheduler).__name__, B, h, w, ctx2.shape[1], latents.device.index, float(guidance_scale), default_lanes(B),
           tuple(scheduler.timesteps.tolist()), tuple(sorted((k, str(v)) for k, v in vars(scheduler.config).items())))
    cache = unet.__dict__.setdefault("_samplers", {})
    smp = cache.get(key)
    if smp is not None and unet._stale(smp._pack_gen):
        cache.clear()
        smp = None
    if smp is None:
        if len(cache) >= 4:
            cache.clear()          # each sampler owns a full set of activation buffers
        smp = cache[key] = CapturedSampler(unet, scheduler, B, h, w, ctx2.shape[1], guidance_scale)
    return smp


# ---------------------------------------------------------------------------------------------------
# StableDiffusionPipeline-compatible wrapper: what finetune_sd.py:517-537 builds and saves,
# what utils.py:181-256 / inference.py:404-429 load, and what inference.py:175-176, 342-351 call.
# UNet, scheduler, text encoder (clip.py) and VAE (vae.py) are b200sd's; objects the caller passes instead (transformers'
# CLIPTextModel, a diffusers AutoencoderKL) are used as they are.  The tokenizer is transformers' CLIPTokenizer (host code).
# ---------------------------------------------------------------------------------------------------
import json
import os
from types import SimpleNamespace

_SCHEDULERS = ("DDIMScheduler", "PNDMScheduler", "DDPMScheduler")


class StableDiffusionPipelineOutput(SimpleNamespace):
    """`.images` (list of PIL images, or a tensor for output_type "pt" / "latent") and `.nsfw_content_detected` (None)."""


class StableDiffusionPipeline:
    """Same constructor keywords, `save_pretrained` / `from_pretrained` directory layout (`model_index.json` + one sub-folder
    per component) and `__call__` signature as diffusers 0.7.2's pipeline.  `safety_checker` / `feature_extractor` are
    accepted and ignored (the reference always passes `safety_checker=None`: inference.py:407, 427; utils.py:190, 225, 250)."""

    config_name = "model_index.json"

    def __init__(self, vae=None, text_encoder=None, tokenizer=None, unet=None, scheduler=None, safety_checker=None,
                 feature_extractor=None):
        if unet is None or scheduler is None:
            raise ValueError("StableDiffusionPipeline needs at least `unet` and `scheduler`")
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.unet, self.scheduler = unet, scheduler
        self.safety_checker, self.feature_extractor = None, feature_extractor
        self._progress_bar_config = {}

    # -- housekeeping the reference touches ----------------------------------------------------------
    @property
    def device(self):
        return self.unet.device

    def to(self, device=None, dtype=None):
        for name in ("unet", "text_encoder", "vae"):
            m = getattr(self, name)
            if m is not None and hasattr(m, "to"):
                setattr(self, name, m.to(device) if dtype is None else m.to(device, dtype=dtype))
        return self

    def set_progress_bar_config(self, **kw):
        self._progress_bar_config = kw

    def enable_attention_slicing(self, *a, **kw):
        """accepted no-op: the fused attention kernel never materialises the S x S scores that slicing exists to bound"""

    disable_attention_slicing = enable_attention_slicing

    @property
    def components(self):
        return dict(vae=self.vae, text_encoder=self.text_encoder, tokenizer=self.tokenizer, unet=self.unet,
                    scheduler=self.scheduler, safety_checker=None, feature_extractor=self.feature_extractor)

    # -- (de)serialisation in the diffusers directory layout -----------------------------------------
    def save_pretrained(self, path, safe_serialization=False):
        os.makedirs(path, exist_ok=True)
        index = {"_class_name": "StableDiffusionPipeline", "_diffusers_version": "0.7.2"}
        for name, obj in self.components.items():
            if obj is None:
                index[name] = [None, None]
                continue
            lib = "diffusers" if name in ("unet", "scheduler", "vae") else "transformers"
            index[name] = [lib, type(obj).__name__]
            sub = os.path.join(path, name)
            if name == "unet":
                obj.save_pretrained(sub, safe_serialization=safe_serialization)
            elif hasattr(obj, "save_pretrained"):
                obj.save_pretrained(sub)
        with open(os.path.join(path, self.config_name), "w") as f:
            json.dump(index, f, indent=2)

    @classmethod
    def from_pretrained(cls, path, torch_dtype=None, safety_checker=None, scheduler=None, **overrides):
        """Loads `unet/` and `scheduler/` with b200sd's classes, and the text encoder / VAE with b200sd's when their
        `text_encoder/` and `vae/` sub-folders exist (b200sd.clip.CLIPTextModel, b200sd.vae.AutoencoderKL); the tokenizer through
        transformers (offline)."""
        from . import schedulers as S
        from .unet import UNet2DConditionModel
        with open(os.path.join(path, cls.config_name)) as f:
            index = json.load(f)
        unet = overrides.pop("unet", None) or UNet2DConditionModel.from_pretrained(path, subfolder="unet", torch_dtype=torch_dtype)
        if scheduler is None:
            cls_name = (index.get("scheduler") or [None, "PNDMScheduler"])[1]
            with open(os.path.join(path, "scheduler", "scheduler_config.json")) as f:
                cls_name = json.load(f).get("_class_name", cls_name)
            if cls_name not in _SCHEDULERS:
                raise ValueError(f"scheduler class {cls_name!r} is not on the reference path ({', '.join(_SCHEDULERS)})")
            kw = {"skip_prk_steps": True} if cls_name == "PNDMScheduler" else {}
            scheduler = getattr(S, cls_name).from_config(path, subfolder="scheduler", **kw)
        tokenizer, text_encoder = overrides.pop("tokenizer", None), overrides.pop("text_encoder", None)
        if tokenizer is None and os.path.isdir(os.path.join(path, "tokenizer")):
            from transformers import CLIPTokenizer
            tokenizer = CLIPTokenizer.from_pretrained(os.path.join(path, "tokenizer"))
        if text_encoder is None and os.path.isdir(os.path.join(path, "text_encoder")):
            from .clip import CLIPTextModel
            text_encoder = CLIPTextModel.from_pretrained(path, subfolder="text_encoder", torch_dtype=torch_dtype)
        vae = overrides.pop("vae", None)
        if vae is None and os.path.isdir(os.path.join(path, "vae")):
            from .vae import AutoencoderKL
            vae = AutoencoderKL.from_pretrained(path, subfolder="vae", torch_dtype=torch_dtype)
        return cls(vae=vae, text_encoder=text_encoder, tokenizer=tokenizer, unet=unet,
                   scheduler=scheduler, safety_checker=None, feature_extractor=overrides.pop("feature_extractor", None))

    # -- sampling ------------------------------------------------------------------------------------
    def _encode_prompt(self, prompt, negative_prompt, device):
        if self.tokenizer is None or self.text_encoder is None:
            raise ValueError("prompt strings need a tokenizer and a text_encoder; pass `prompt_embeds` (2B, 77, 768) instead")
        tok = lambda texts: self.tokenizer(texts, padding="max_length", max_length=self.tokenizer.model_max_length,
                                           truncation=True, return_tensors="pt").input_ids.to(device)
        cond = self.text_encoder(tok(prompt))[0]
        uncond = self.text_encoder(tok(negative_prompt if negative_prompt is not None else [""] * len(prompt)))[0]
        return torch.cat([uncond, cond])          # App. B.4: row block 0 = unconditional

    @torch.no_grad()
    def __call__(self, prompt=None, height=512, width=512, num_inference_steps=50, guidance_scale=7.5, negative_prompt=None,
                 eta=0.0, generator=None, latents=None, output_type="pil", return_dict=True, prompt_embeds=None, **kw):
        """inference.py:175-176, 342-351: `pipeline(prompts, height=, width=, num_inference_steps=50, guidance_scale=7.5,
        latents=...)`.  `prompt_embeds` = cat([uncond, cond]) lets a caller without CLIP drive the hot path directly."""
        if eta != 0.0:
            raise NotImplementedError("eta != 0 is not on the reference path")
        if height % 8 or width % 8:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        device = self.device
        if prompt_embeds is None:
            prompt = [prompt] if isinstance(prompt, str) else list(prompt)
            ctx2 = self._encode_prompt(prompt, [negative_prompt] * len(prompt) if isinstance(negative_prompt, str) else negative_prompt,
                                       device)
        else:
            ctx2 = prompt_embeds.to(device)
        B = ctx2.shape[0] // 2
        shape = (B, self.unet.in_channels, height // 8, width // 8)
        if latents is None:
            latents = torch.randn(shape, generator=generator, device=device if generator is None or generator.device.type != "cpu"
                                  else "cpu").to(device)
        elif tuple(latents.shape) != shape:
            raise ValueError(f"Unexpected latents shape, got {tuple(latents.shape)}, expected {shape}")
        lat = denoise_loop(self.unet, self.scheduler, latents.to(device).float(), ctx2.float(), num_inference_steps, guidance_scale)
        if output_type == "latent" or self.vae is None:
            images = lat
        else:
            dec = self.vae.decode(lat.to(next(self.vae.parameters()).dtype) / 0.18215)
            img = (getattr(dec, "sample", dec) / 2 + 0.5).clamp(0, 1)
            if output_type == "pt":
                images = img
            else:
                arr = (img.permute(0, 2, 3, 1).float().cpu().numpy() * 255).round().astype("uint8")
                from PIL import Image
                images = [Image.fromarray(a) for a in arr]
        out = StableDiffusionPipelineOutput(images=images, nsfw_content_detected=None)
        return out if return_dict else (images, None)
