"""Generates tests/golden/sd15_kat.json: closed-form known answers of SURVEY.md App. B.5 (scheduler
tables, timestep lists, timestep embedding) plus small seeded input/output vectors of the ORACLE's
scheduler steps.  The reference has no fixtures of its own and diffusers is not importable here
(parity unpinned), so these pin the oracle against accidental change, not against diffusers.
    python tests/golden/make_golden.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import schedulers_ref as R
from oracle.unet_ref import timestep_embedding

out = {}
s = R.DDIMSchedulerRef(clip_sample=False, set_alpha_to_one=False)
out["alphas_cumprod"] = {str(i): float(s.alphas_cumprod[i]) for i in (0, 1, 20, 500, 980, 981, 999)}
s.set_timesteps(50)
out["ddim_timesteps_50"] = s.timesteps.tolist()
p = R.PNDMSchedulerRef(skip_prk_steps=True, steps_offset=1)
p.set_timesteps(50)
out["plms_timesteps_50_offset1"] = p.timesteps.tolist()
e = timestep_embedding(torch.tensor([980]))
out["temb_980_cos_0_3"] = e[0, :3].tolist()
out["temb_980_sin_0_3"] = e[0, 160:163].tolist()
g = torch.Generator().manual_seed(7)
x = torch.randn(1, 4, 4, 4, generator=g)
eps = torch.randn(1, 4, 4, 4, generator=g)
out["ddim_step_t500"] = {"x": x.flatten().tolist(), "eps": eps.flatten().tolist(),
                         "prev": s.step(eps, 500, x).prev_sample.flatten().tolist()}
t = torch.tensor([333])
out["add_noise_t333"] = R.DDPMSchedulerRef().add_noise(x, eps, t).flatten().tolist()
xs, cur = [], x.clone()
for i, tt in enumerate(p.timesteps[:6]):
    ee = torch.randn(1, 4, 4, 4, generator=torch.Generator().manual_seed(50 + i))
    cur = p.step(ee, tt, cur).prev_sample
    xs.append(cur.flatten().tolist())
out["plms_first6"] = xs
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sd15_kat.json"), "w") as f:
    json.dump(out, f)
print("wrote sd15_kat.json")
