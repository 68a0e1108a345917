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
