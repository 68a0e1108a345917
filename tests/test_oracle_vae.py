"""CPU self-checks of the AutoencoderKL oracle (oracle/vae_ref.py; its layers are pinned to diffusers' own known answers in tests/test_oracle_diffusers_kat.py, its wiring is checked here): the
public SD v1.x VAE's parameter / tensor counts and key names, explicit-softmax attention vs torch SDPA, the asymmetric
downsample padding vs an explicit unfold, posterior arithmetic, and the product module's state-dict surface."""
import pytest
import torch
import torch.nn.functional as F

from oracle.vae_ref import TINY_VAE_OVERRIDES, AttentionBlock, AutoencoderKLRef, DiagonalGaussianDistribution, _Down, make_oracle_vae


def test_sd15_vae_parameter_count_and_keys():
    with torch.device("meta"):
        m = AutoencoderKLRef()
    assert sum(p.numel() for p in m.parameters()) == 83_653_863
    sd = m.state_dict()
    assert len(sd) == 248
    for k, shape in (("encoder.conv_in.weight", (128, 3, 3, 3)), ("encoder.down_blocks.1.resnets.0.conv_shortcut.weight", (256, 128, 1, 1)),
                     ("encoder.down_blocks.2.downsamplers.0.conv.weight", (512, 512, 3, 3)),
                     ("encoder.mid_block.attentions.0.query.weight", (512, 512)), ("encoder.conv_out.weight", (8, 512, 3, 3)),
                     ("quant_conv.weight", (8, 8, 1, 1)), ("post_quant_conv.bias", (4,)), ("decoder.conv_in.weight", (512, 4, 3, 3)),
                     ("decoder.up_blocks.2.resnets.0.conv_shortcut.weight", (256, 512, 1, 1)),
                     ("decoder.up_blocks.0.upsamplers.0.conv.weight", (512, 512, 3, 3)), ("decoder.up_blocks.3.resnets.2.norm2.bias", (128,)),
                     ("decoder.mid_block.attentions.0.proj_attn.bias", (512,)), ("decoder.conv_out.weight", (3, 128, 3, 3))):
        assert tuple(sd[k].shape) == shape, k
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in sd and "decoder.up_blocks.3.upsamplers.0.conv.weight" not in sd


def test_shapes_and_scale_factor():
    m = make_oracle_vae(0, **TINY_VAE_OVERRIDES)
    x = torch.randn(2, 3, 64, 32)
    with torch.no_grad():
        post = m.encode(x).latent_dist
        assert tuple(post.mean.shape) == (2, 4, 8, 4)
        assert tuple(m.decode(post.mode()).sample.shape) == (2, 3, 64, 32)


def test_attention_block_equals_sdpa():
    torch.manual_seed(0)
    a = AttentionBlock(64, 32).eval()
    x = torch.randn(2, 64, 6, 5)
    with torch.no_grad():
        t = a.group_norm(x).view(2, 64, 30).transpose(1, 2)
        want = a.proj_attn(F.scaled_dot_product_attention(a.query(t)[:, None], a.key(t)[:, None], a.value(t)[:, None])[:, 0])
        want = want.transpose(1, 2).reshape(2, 64, 6, 5) + x
        assert float((a(x) - want).abs().max()) <= 1e-5


def test_downsample_pads_right_and_bottom_only():
    torch.manual_seed(0)
    d = _Down(4).eval()
    x = torch.randn(1, 4, 6, 6)
    with torch.no_grad():
        got = d(x)
        cols = F.unfold(F.pad(x, (0, 1, 0, 1)), 3, stride=2)                      # window (oy, ox) starts at (2 oy, 2 ox)
        want = (d.conv.weight.view(4, -1) @ cols + d.conv.bias[None, :, None]).view(1, 4, 3, 3)
    assert tuple(got.shape) == (1, 4, 3, 3) and float((got - want).abs().max()) <= 1e-5
    assert float(got[0, :, 0, 0].sub((d.conv.weight * x[0, :, :3, :3]).sum((1, 2, 3)) + d.conv.bias).abs().max()) <= 1e-5


def test_posterior_arithmetic():
    mom = torch.randn(2, 8, 4, 4)
    mom[0, 4] = 100.0
    p = DiagonalGaussianDistribution(mom)
    assert float(p.logvar.max()) == 20.0 and torch.equal(p.mode(), mom[:, :4])
    n = torch.randn(2, 4, 4, 4)
    assert torch.allclose(p.sample(noise=n), mom[:, :4] + torch.exp(0.5 * p.logvar) * n)


def test_product_module_has_the_diffusers_state_dict_surface(tmp_path):
    from b200sd.vae import AutoencoderKL
    o = make_oracle_vae(1, **TINY_VAE_OVERRIDES)
    m = AutoencoderKL(**TINY_VAE_OVERRIDES)
    m.load_state_dict(o.state_dict(), strict=True)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in o.state_dict().items()}
    with torch.device("meta"):
        full = AutoencoderKL()
    assert sum(p.numel() for p in full.parameters()) == 83_653_863
    d = str(tmp_path / "vae")
    m.save_pretrained(d)
    m2 = AutoencoderKL.from_pretrained(d)
    assert all(torch.equal(v, m2.state_dict()[k]) for k, v in m.state_dict().items())
    with pytest.raises(Exception):
        m.decode(torch.randn(1, 4, 8, 8))            # CPU tensors: no fallback


def test_quant_conv_is_folded_into_the_encoder_conv_out_exactly():
    """vae.py packs `quant_conv(conv_out(x))` as ONE conv (W' = Wq . Wc per tap, b' = Wq bc + bq) with hi/lo-split bf16 weights
    padded to a 32-wide tensor-core tile: the packed weights, put back together, must reproduce the two-layer result."""
    from b200sd.vae import AutoencoderKL
    o = make_oracle_vae(3, **TINY_VAE_OVERRIDES)
    m = AutoencoderKL(**TINY_VAE_OVERRIDES)
    m.load_state_dict(o.state_dict(), strict=True)
    m._pack_weights()                                    # pure tensor reshuffles: runs on CPU
    pk = m._packed["enc.conv_out"]
    cin = o.encoder.conv_out.weight.shape[1]
    w = pk["w_tc"].float().view(32, 9, 2 * cin)
    assert float(w[8:].abs().max()) == 0.0               # rows beyond the 8 moment channels are padding
    w_fold = (w[:8, :, :cin] + w[:8, :, cin:]).view(8, 3, 3, cin).permute(0, 3, 1, 2)      # hi + lo, back to OIHW
    x = torch.randn(2, cin, 6, 5)
    with torch.no_grad():
        want = o.quant_conv(o.encoder.conv_out(x))
        got = F.conv2d(x, w_fold, pk["b"], padding=1)
    assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())
    # decoder conv_out: same hi/lo packing of the 3 image channels; post_quant_conv kept as a [4][4] matrix
    wd = m._packed["dec.conv_out"]["w_tc"].float()
    c0 = o.decoder.conv_out.weight.shape[1]
    wd = wd.view(32, 9, 2 * c0)
    back = (wd[:3, :, :c0] + wd[:3, :, c0:]).view(3, 3, 3, c0).permute(0, 3, 1, 2)
    assert float((back - o.decoder.conv_out.weight.detach()).abs().max()) <= 1e-5 * float(o.decoder.conv_out.weight.abs().max())
    assert torch.equal(m._packed["post_quant"]["w"], o.post_quant_conv.weight.detach().view(4, 4))
    # the 3-channel encoder conv_in is packed with a zero fourth input channel
    wi = m._packed["enc.conv_in"]["w"].view(-1, 9, 4)
    assert float(wi[:, :, 3].abs().max()) == 0.0
