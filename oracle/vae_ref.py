"""ORACLE (test infrastructure, not product code) -- SD v1.x AutoencoderKL restated in plain fp32 PyTorch.

PARITY: LAYERS PINNED, BLOCK WIRING UNPINNED.  The arithmetic lives in the un-vendored third-party package `diffusers==0.7.2`
(pinned at /root/reference/env.yaml:112; models/vae.py, unet_2d_blocks.py, resnet.py, attention.py) which is neither installed
nor installable here, and the reference holds no fixtures.  This module restates the published 0.7.2 algorithm.  Its layers are
pinned against the known answers diffusers 0.7.2 holds in its own test suite (tests/golden/diffusers_0_7_2_kat.json, checked by
tests/test_oracle_diffusers_kat.py): `AttentionBlock` at the SD shape (512 channels, one head) and the decoder upsampler directly,
`ResnetBlock2D` through the UNet oracle's pinned block (same weights, time-embedding projection zeroed), the encoder / decoder
blocks against DownEncoderBlock2D / UpDecoderBlock2D vectors (diffusers tests/test_unet_blocks.py).  The encoder/decoder stack
wiring, the asymmetric stride-2 padding and the latent distribution have no checkpoint-free known answer in diffusers and stay
anchored on the reference's call sites:

  * `AutoencoderKL.from_pretrained(path, subfolder="vae")`                                   -- finetune_sd.py:325-327
  * `latents = vae.encode(batch["pixel_values"]).latent_dist.sample() * 0.18215`           -- finetune_sd.py:460-462
  * `vae.decode(latents / 0.18215).sample` inside every `pipeline(...)` call                 -- inference.py:175-176, 342-351

Self-checks for the unpinned part (tests/test_oracle_vae.py): 83 653 863 parameters (the public SD v1.x VAE),
248 state-dict tensors with the diffusers 0.7.2 key names (`encoder.down_blocks.0.resnets.0.norm1.weight`,
`decoder.mid_block.attentions.0.query.weight`, `quant_conv.weight`, ...), explicit-softmax attention vs torch SDPA, the
asymmetric (0,1,0,1) downsample padding vs an explicit unfold.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this file; the product (package `b200sd`) never does.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

SD15_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                       layers_per_block=2, norm_num_groups=32, act_fn="silu", sample_size=512,
                       down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4)
TINY_VAE_OVERRIDES = dict(block_out_channels=(64, 64, 128, 128))


class ResnetBlock2D(nn.Module):
    """resnet.py ResnetBlock2D with temb_channels=None, eps=1e-6, swish, output_scale_factor=1"""

    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.conv_shortcut = nn.Conv2d(cin, cout, 1)          # use_nin_shortcut

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if hasattr(self, "conv_shortcut"):
            x = self.conv_shortcut(x)
        return x + h


class AttentionBlock(nn.Module):
    """attention.py AttentionBlock, num_head_channels=None (one head over all channels), rescale_output_factor=1"""

    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.query, self.key, self.value, self.proj_attn = (nn.Linear(c, c) for _ in range(4))

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.query(t), self.key(t), self.value(t)
        scale = 1.0 / math.sqrt(math.sqrt(c))                       # applied to q AND k
        p = torch.softmax((q * scale) @ (k * scale).transpose(-1, -2), dim=-1)
        o = self.proj_attn(p @ v)
        return o.transpose(1, 2).reshape(b, c, h, w) + x


class _Down(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))                     # Downsample2D(padding=0): pad right / bottom only


class _Up(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _EncBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, down):
        super().__init__()
        self.resnets = nn.ModuleList(ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers))
        if down:
            self.downsamplers = nn.ModuleList([_Down(cout)])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.downsamplers[0](x) if hasattr(self, "downsamplers") else x


class _DecBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, up):
        super().__init__()
        self.resnets = nn.ModuleList(ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers))
        if up:
            self.upsamplers = nn.ModuleList([_Up(cout)])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.upsamplers[0](x) if hasattr(self, "upsamplers") else x


class _Mid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([AttentionBlock(c, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, groups), ResnetBlock2D(c, c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList(_EncBlock(boc[max(i - 1, 0)], boc[i], cfg.layers_per_block, g, i < len(boc) - 1)
                                         for i in range(len(boc)))
        self.mid_block = _Mid(boc[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(self.mid_block(x))))


class Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        rev = list(reversed(boc))
        self.conv_in = nn.Conv2d(cfg.latent_channels, rev[0], 3, padding=1)
        self.mid_block = _Mid(rev[0], g)
        self.up_blocks = nn.ModuleList(_DecBlock(rev[max(i - 1, 0)], rev[i], cfg.layers_per_block + 1, g, i < len(rev) - 1)
                                       for i in range(len(rev)))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    """vae.py: moments = [mean | logvar] along channels, logvar clamped to [-30, 20]"""

    def __init__(self, parameters):
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std, self.var = torch.exp(0.5 * self.logvar), torch.exp(self.logvar)

    def sample(self, generator=None, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean

    def kl(self):
        return 0.5 * torch.sum(self.mean ** 2 + self.var - 1.0 - self.logvar, dim=[1, 2, 3])


class AutoencoderKLRef(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        self.config = SimpleNamespace(**{**SD15_VAE_CONFIG, **overrides})
        self.encoder, self.decoder = Encoder(self.config), Decoder(self.config)
        lc = self.config.latent_channels
        self.quant_conv = nn.Conv2d(2 * lc, 2 * lc, 1)
        self.post_quant_conv = nn.Conv2d(lc, lc, 1)

    def encode(self, x):
        return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(self.quant_conv(self.encoder(x))))

    def decode(self, z):
        return SimpleNamespace(sample=self.decoder(self.post_quant_conv(z)))

    def forward(self, sample, sample_posterior=False, generator=None):
        post = self.encode(sample).latent_dist
        return self.decode(post.sample(generator) if sample_posterior else post.mode())


def make_oracle_vae(seed=0, **overrides):
    """torch default init under the seed; GroupNorm affines perturbed, attention q/k sharpened (x2) so that they are exercised"""
    torch.manual_seed(seed)
    m = AutoencoderKLRef(**overrides)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
            if n.endswith("query.weight") or n.endswith("key.weight"):
                p.mul_(2.0)
    return m.eval()
