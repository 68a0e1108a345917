"""Generates tests/golden/clip_text_kat.json from `transformers.CLIPTextModel` ITSELF -- the library class the reference
instantiates (finetune_sd.py:322-324; transformers is pinned at 4.29.2 in env.yaml:128, this image has 5.5.0: same published
algorithm).  Weights: oracle.clip_ref.make_oracle_clip(seed=0, **TINY_CLIP_OVERRIDES).state_dict() loaded into the
transformers model (keys are identical); inputs: seeded token ids.  Stored: the ids, the full last_hidden_state of prompt 0
(77 x 128), per-prompt checksums, and gradient checksums of sum(last_hidden_state * w) for a seeded w.
    python tests/golden/make_clip_golden.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import transformers
from transformers import CLIPTextConfig, CLIPTextModel
from oracle.clip_ref import TINY_CLIP_OVERRIDES, make_oracle_clip

sd = make_oracle_clip(seed=0, **TINY_CLIP_OVERRIDES).state_dict()
cfg = CLIPTextConfig(hidden_act="quick_gelu", max_position_embeddings=77, eos_token_id=999, bos_token_id=0, pad_token_id=1,
                     **TINY_CLIP_OVERRIDES)
hf = CLIPTextModel(cfg).eval()
hf.load_state_dict(sd, strict=True)
g = torch.Generator().manual_seed(11)
ids = torch.randint(2, 998, (3, 77), generator=g)
ids[:, 0], ids[:, -1] = 0, 999
w = torch.randn(3, 77, 128, generator=g)
out = hf(ids)
last = out[0]
(last * w).sum().backward()
grads = {n: p.grad for n, p in hf.named_parameters()}
pick = ["text_model.embeddings.position_embedding.weight", "text_model.encoder.layers.0.self_attn.q_proj.weight",
        "text_model.encoder.layers.0.self_attn.v_proj.bias", "text_model.encoder.layers.1.mlp.fc1.weight",
        "text_model.encoder.layers.1.layer_norm2.weight", "text_model.final_layer_norm.bias"]
gold = {"transformers_version": transformers.__version__, "ids": ids.tolist(),
        "last_hidden_state_prompt0": last[0].detach().flatten().tolist(),
        "last_hidden_state_sum": [float(last[b].sum()) for b in range(3)],
        "last_hidden_state_abs_sum": [float(last[b].abs().sum()) for b in range(3)],
        "pooled_prompt0_first8": out[1][0, :8].detach().tolist(),
        "grad_abs_sum": {n: float(grads[n].abs().sum()) for n in pick},
        "grad_first4": {n: grads[n].flatten()[:4].tolist() for n in pick}}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "clip_text_kat.json"), "w") as f:
    json.dump(gold, f)
print("wrote clip_text_kat.json")
