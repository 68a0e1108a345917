"""CPU tests that PIN the CLIP text-encoder oracle (oracle/clip_ref.py): against `transformers.CLIPTextModel` itself -- the
library class the reference instantiates (finetune_sd.py:322-324) -- forward and autograd gradients on identical weights, and
against the committed vectors that tests/golden/make_clip_golden.py generated from the transformers model."""
import json
import os

import pytest
import torch

from oracle.clip_ref import SD15_CLIP_CONFIG, TINY_CLIP_OVERRIDES, CLIPTextModelRef, make_oracle_clip

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "clip_text_kat.json")))


def _hf(overrides, eos):
    from transformers import CLIPTextConfig, CLIPTextModel
    cfg = CLIPTextConfig(hidden_act="quick_gelu", max_position_embeddings=77, eos_token_id=eos, bos_token_id=0, pad_token_id=1,
                         **overrides)
    return CLIPTextModel(cfg).eval()


def test_sd15_text_encoder_parameter_count_and_keys():
    with torch.device("meta"):
        m = CLIPTextModelRef()
    assert sum(p.numel() for p in m.parameters()) == 123_060_480
    sd = m.state_dict()
    assert len(sd) == 196
    for k in ("text_model.embeddings.token_embedding.weight", "text_model.embeddings.position_embedding.weight",
              "text_model.encoder.layers.11.self_attn.out_proj.bias", "text_model.encoder.layers.0.mlp.fc1.weight",
              "text_model.encoder.layers.5.layer_norm2.weight", "text_model.final_layer_norm.weight"):
        assert k in sd
    assert tuple(sd["text_model.embeddings.token_embedding.weight"].shape) == (49408, 768)
    assert SD15_CLIP_CONFIG["max_position_embeddings"] == 77


def test_oracle_matches_golden_vectors_generated_from_transformers():
    o = make_oracle_clip(seed=0, **TINY_CLIP_OVERRIDES)
    ids = torch.tensor(GOLD["ids"])
    w = torch.randn(3, 77, 128, generator=_gen_after_ids())
    for p in o.parameters():
        p.requires_grad_(True)
    last, pooled = o(ids)
    want0 = torch.tensor(GOLD["last_hidden_state_prompt0"]).view(77, 128)
    assert float((last[0].detach() - want0).abs().max()) <= 2e-5 * float(want0.abs().max())
    for b in range(3):
        assert float(last[b].detach().sum()) == pytest.approx(GOLD["last_hidden_state_sum"][b], rel=1e-4, abs=1e-3)
        assert float(last[b].detach().abs().sum()) == pytest.approx(GOLD["last_hidden_state_abs_sum"][b], rel=1e-5)
    assert pooled[0, :8].detach().tolist() == pytest.approx(GOLD["pooled_prompt0_first8"], rel=1e-4, abs=1e-5)
    (last * w).sum().backward()
    grads = dict(o.named_parameters())
    for n, s in GOLD["grad_abs_sum"].items():
        assert float(grads[n].grad.abs().sum()) == pytest.approx(s, rel=1e-4), n
        assert grads[n].grad.flatten()[:4].tolist() == pytest.approx(GOLD["grad_first4"][n], rel=1e-3, abs=1e-5), n


def _gen_after_ids():
    """the generator state make_clip_golden.py had when it drew w: seed 11, after the ids"""
    g = torch.Generator().manual_seed(11)
    torch.randint(2, 998, (3, 77), generator=g)
    return g


@pytest.mark.parametrize("overrides", [TINY_CLIP_OVERRIDES, dict(vocab_size=2000, num_hidden_layers=1)])
def test_oracle_matches_transformers_forward_and_gradients(overrides):
    """Live check against the installed transformers (tiny config; and the real width -- 768 / 12 heads / 3072 -- with one layer
    and a small vocabulary so it stays fast)."""
    pytest.importorskip("transformers")
    o = make_oracle_clip(seed=1, **overrides)
    full = dict(hidden_size=o.config.hidden_size, intermediate_size=o.config.intermediate_size,
                num_attention_heads=o.config.num_attention_heads, num_hidden_layers=o.config.num_hidden_layers,
                vocab_size=o.config.vocab_size)
    hf = _hf(full, eos=overrides["vocab_size"] - 1)
    assert set(hf.state_dict().keys()) == set(o.state_dict().keys())
    hf.load_state_dict(o.state_dict(), strict=True)
    g = torch.Generator().manual_seed(5)
    V, C = overrides["vocab_size"], o.config.hidden_size
    ids = torch.randint(2, V - 2, (2, 77), generator=g)
    ids[:, -1] = V - 1
    w = torch.randn(2, 77, C, generator=g)
    for m in (o, hf):
        for p in m.parameters():
            p.requires_grad_(True)
    a = hf(ids)
    b = o(ids)
    assert float((a[0] - b[0]).abs().max()) <= 1e-5 * float(a[0].abs().max())
    assert float((a[1] - b[1]).abs().max()) <= 1e-5 * float(a[1].abs().max())
    (a[0] * w).sum().backward()
    (b[0] * w).sum().backward()
    go = dict(o.named_parameters())
    floor = 1e-6 * max(float(p.grad.abs().max()) for p in hf.parameters())   # k_proj.bias gradients are mathematically zero
    for n, p in hf.named_parameters():                                        # (softmax is shift-invariant): pure rounding noise
        ref = p.grad
        assert float((go[n].grad - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + floor, n


def test_product_module_has_the_transformers_state_dict_and_config_surface(tmp_path):
    """b200sd.clip.CLIPTextModel: same keys / shapes as transformers, save_pretrained -> from_pretrained round trip, and the
    saved directory loads into transformers' own class (CPU: no kernels involved)."""
    from b200sd.clip import CLIPTextModel
    o = make_oracle_clip(seed=2, **TINY_CLIP_OVERRIDES)
    m = CLIPTextModel(**TINY_CLIP_OVERRIDES)
    m.load_state_dict(o.state_dict(), strict=True)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in o.state_dict().items()}
    for safe in (False, True):
        d = str(tmp_path / f"te{int(safe)}")
        m.save_pretrained(d, safe_serialization=safe)
        m2 = CLIPTextModel.from_pretrained(d)
        for k, v in m.state_dict().items():
            assert torch.equal(v, m2.state_dict()[k])
        assert m2.config.hidden_size == 128 and m2.config.num_hidden_layers == 2
    transformers = pytest.importorskip("transformers")
    hf = transformers.CLIPTextModel.from_pretrained(str(tmp_path / "te1"))
    ids = torch.randint(2, 900, (1, 77))
    with torch.no_grad():
        assert float((hf(ids)[0] - o(ids)[0]).abs().max()) <= 1e-4
    with pytest.raises(Exception):
        m(ids)           # CPU tensors: no fallback


def test_loads_a_directory_written_by_transformers_itself(tmp_path):
    """`CLIPTextModel.from_pretrained(path, subfolder="text_encoder")` (finetune_sd.py:322-324) on a checkpoint saved by the
    library class: config.json keys and model.safetensors tensor names are read as they are."""
    transformers = pytest.importorskip("transformers")
    from b200sd.clip import CLIPTextModel
    cfg = transformers.CLIPTextConfig(hidden_act="quick_gelu", max_position_embeddings=77, eos_token_id=999, bos_token_id=0,
                                      pad_token_id=1, **TINY_CLIP_OVERRIDES)
    torch.manual_seed(4)
    hf = transformers.CLIPTextModel(cfg).eval()
    hf.save_pretrained(str(tmp_path / "text_encoder"))
    ours = CLIPTextModel.from_pretrained(str(tmp_path), subfolder="text_encoder")
    assert ours.config.hidden_size == 128 and ours.config.num_hidden_layers == 2 and ours.config.vocab_size == 1000
    sd_hf = {k: v for k, v in hf.state_dict().items() if not k.endswith("position_ids")}
    assert set(sd_hf) == set(ours.state_dict())
    assert all(torch.equal(v, ours.state_dict()[k]) for k, v in sd_hf.items())
