"""ORACLE (test infrastructure, not product code) -- bitsandbytes 0.35.4 `AdamW8bit` (block-wise 8-bit optimizer state)
restated in plain fp32 PyTorch on the CPU.

PARITY UNPINNED.  The reference's default optimizer is `bnb.optim.AdamW8bit(params, lr, weight_decay, min_8bit_size=16384)`
(finetune_sd.py:300 `use_8bit_adam=True`, :407-420); `bitsandbytes==0.35.4` (env.yaml:111) is un-vendored, not installed and not
installable here, and neither the reference nor this image holds a test vector for it.  This module restates the published
algorithm:

  * `create_dynamic_map` (bitsandbytes/functional.py): the 256-entry "dynamic tree" code book -- signed for the first moment,
    unsigned for the second; 7 decades, decade i holding 2^i (signed) / 2^(i+1) (unsigned) linearly spaced fractions, plus 0 and 1;
  * `kOptimizerStatic8bit2StateBlockwise<ADAM>` (csrc/kernels.cu): per block of 2048 values, de-quantise both moments with the
    block's absmax, Adam moment update, new absmax = max |moment| over the block, parameter update
    `p += step_size * m / (sqrt(v) + correction2 * eps)` with `step_size = -lr * correction2 / correction1`,
    `correction2 = sqrt(1 - beta2^t)`, decoupled decay `p *= 1 - lr * wd` AFTER the update, re-quantise to the nearest code of
    `moment / new absmax`, and keep the first moment's sign ("make sure state1 term has still the same sign after quantization");
  * `kOptimizer32bit2State<ADAM>`: the same update with fp32 moments for tensors below `min_8bit_size`.

Where this restatement knowingly differs from bitsandbytes: (1) the nearest code is found exactly (midpoint rule, ties to the
lower code), bitsandbytes' `quantize_2D` binary search can land one code off near a midpoint; (2) blocks of 2048 run over the
FLAT parameter buffer of the model (b200sd.train.FlatParams), not over each tensor separately, so a block can hold the tail of one
tensor and the head of the next; (3) every product / sum is a separately rounded fp32 operation in the order written below (the
CUDA kernel uses the `__f*_rn` intrinsics in the same order), which makes kernel-vs-oracle parity BIT-EXACT for the codes, the
absmax tables and the parameters.  The self-checks standing in for the missing pin are in tests/test_oracle_adam8bit.py (code-book
structure, quantisation error bounds, agreement with fp32 torch.optim.AdamW over a training-like gradient sequence).

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this file; the product (package `b200sd`) never does.
"""
from __future__ import annotations

import math

import numpy as np
import torch

BLOCK = 2048          # bitsandbytes' block size of the 8-bit optimizers
CHUNK = 64            # alignment of every parameter region in the flat buffer (b200sd.train._ALIGN)
MODE_8BIT, MODE_SKIP = -1, -2


def create_dynamic_map(signed: bool = True, n: int = 7) -> torch.Tensor:
    """bitsandbytes/functional.py create_dynamic_map (0.35.x): sorted fp32 tensor of 256 values in [-1, 1] (signed) / [0, 1]."""
    data = []
    additional_items = 2 ** (7 - n) - 1
    if not signed:
        additional_items = 2 * additional_items
    for i in range(n):
        fraction_items = 2 ** (i + 7 - n) + 1 if signed else 2 ** (i + 7 - n + 1) + 1
        boundaries = torch.linspace(0.1, 1, fraction_items)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(n - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(n - 1) + i)) * means).tolist()
    if additional_items > 0:
        boundaries = torch.linspace(0.1, 1, additional_items + 1)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(n - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(n - 1) + i)) * means).tolist()
    data.append(0)
    data.append(1.0)
    data.sort()
    return torch.tensor(data, dtype=torch.float32)


def sqrt_rn(x: torch.Tensor) -> torch.Tensor:
    """correctly rounded fp32 square root (numpy's; torch.sqrt on the CPU goes through a vector math library that is 1 ulp off for
    ~0.7 % of its arguments, which is enough to break bit-exact parity with the kernel's __fsqrt_rn)"""
    return torch.from_numpy(np.sqrt(x.detach().numpy()))


def midpoints(qmap: torch.Tensor) -> torch.Tensor:
    return (qmap[:-1] + qmap[1:]) * 0.5


def quantize_nearest(x: torch.Tensor, qmap: torch.Tensor) -> torch.Tensor:
    """code = number of midpoints strictly below x (nearest code, ties to the lower one)"""
    return torch.bucketize(x, midpoints(qmap), right=False).to(torch.uint8)


def quantize_blockwise(x: torch.Tensor, qmap: torch.Tensor):
    """reference-style round trip helper: (codes, absmax per block of BLOCK values)"""
    n = x.numel()
    nb = (n + BLOCK - 1) // BLOCK
    xp = torch.zeros(nb * BLOCK, dtype=torch.float32)
    xp[:n] = x.flatten()
    absmax = xp.view(nb, BLOCK).abs().amax(dim=1)
    scaled = torch.where(absmax[:, None] > 0, xp.view(nb, BLOCK) / absmax[:, None], torch.zeros(()))
    return quantize_nearest(scaled.flatten(), qmap)[:n], absmax


def dequantize_blockwise(codes: torch.Tensor, absmax: torch.Tensor, qmap: torch.Tensor) -> torch.Tensor:
    idx = torch.arange(codes.numel()) // BLOCK
    return qmap[codes.long()] * absmax[idx]


def step_constants(lr, beta1, beta2, eps, weight_decay, step):
    """host-side scalars exactly as the kernel launcher forms them: the hyper-parameters arrive as C floats (rounded to fp32
    first), the derived constants are computed in double and rounded to fp32 once"""
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float64).to(torch.float32))
    lr, beta1, beta2, eps, weight_decay = f32(lr), f32(beta1), f32(beta2), f32(eps), f32(weight_decay)
    c1 = 1.0 - math.pow(beta1, step)
    c2 = math.sqrt(1.0 - math.pow(beta2, step))
    return dict(beta1=beta1, beta2=beta2, omb1=f32(1.0 - beta1), omb2=f32(1.0 - beta2), eps_c2=f32(eps * c2),
                step_size=f32(-lr * c2 / c1), decay=f32(1.0 - lr * weight_decay), apply_decay=weight_decay > 0.0)


class AdamW8bitRef:
    """State and one step over a flat fp32 parameter buffer.  `chunk_mode[k]` describes elements [64k, 64k + 64):
    MODE_8BIT, MODE_SKIP (frozen / padding: untouched) or an offset >= 0 into the compact fp32 moments of the small tensors."""

    def __init__(self, n: int, chunk_mode: torch.Tensor | None = None, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        assert n % CHUNK == 0
        self.n, self.lr, self.betas, self.eps, self.weight_decay = n, lr, betas, eps, weight_decay
        self.chunk_mode = torch.full((n // CHUNK,), MODE_8BIT, dtype=torch.int32) if chunk_mode is None else chunk_mode.to(torch.int32)
        self.nblocks = (n + BLOCK - 1) // BLOCK
        self.qmap1, self.qmap2 = create_dynamic_map(True), create_dynamic_map(False)
        self.state1 = torch.zeros(n, dtype=torch.uint8)
        self.state2 = torch.zeros(n, dtype=torch.uint8)
        self.absmax1 = torch.zeros(self.nblocks, dtype=torch.float32)
        self.absmax2 = torch.zeros(self.nblocks, dtype=torch.float32)
        n_small = int((self.chunk_mode >= 0).sum()) * CHUNK
        self.small_m = torch.zeros(n_small, dtype=torch.float32)
        self.small_v = torch.zeros(n_small, dtype=torch.float32)
        self.steps = 0

    def _element_modes(self):
        mode = self.chunk_mode.repeat_interleave(CHUNK)
        within = torch.arange(self.n, dtype=torch.int64) % CHUNK
        return mode, within

    def step(self, param: torch.Tensor, grad: torch.Tensor, grad_scale: float = 1.0, zero_grad: bool = True):
        """in place on `param` / `grad` (fp32, n elements); returns the bf16 copy of the updated parameters"""
        self.steps += 1
        k = step_constants(self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.steps)
        t = lambda v: torch.tensor(v, dtype=torch.float32)
        mode, within = self._element_modes()
        is8, small, live = mode == MODE_8BIT, mode >= 0, mode != MODE_SKIP
        blk = torch.arange(self.n) // BLOCK
        small_idx = (mode.long() + within)[small]
        # de-quantise / load the moments
        s1 = torch.zeros(self.n)
        s2 = torch.zeros(self.n)
        s1[is8] = (self.qmap1[self.state1.long()] * self.absmax1[blk])[is8]
        s2[is8] = (self.qmap2[self.state2.long()] * self.absmax2[blk])[is8]
        s1[small] = self.small_m[small_idx]
        s2[small] = self.small_v[small_idx]
        g = grad * t(grad_scale)
        s2 = s2 * t(k["beta2"]) + (t(k["omb2"]) * g) * g
        s1 = s1 * t(k["beta1"]) + t(k["omb1"]) * g
        upd = s1 / (sqrt_rn(s2) + t(k["eps_c2"]))
        newp = param + t(k["step_size"]) * upd
        if k["apply_decay"]:
            newp = newp * t(k["decay"])
        param[live] = newp[live]
        if zero_grad:
            grad[live] = 0.0
        self.small_m[small_idx] = s1[small]
        self.small_v[small_idx] = s2[small]
        # new block absmax over the 8-bit elements only, then re-quantise
        pad = self.nblocks * BLOCK - self.n
        a1 = torch.cat([torch.where(is8, s1.abs(), torch.zeros(())), torch.zeros(pad)]).view(self.nblocks, BLOCK).amax(dim=1)
        a2 = torch.cat([torch.where(is8, s2.abs(), torch.zeros(())), torch.zeros(pad)]).view(self.nblocks, BLOCK).amax(dim=1)
        # scaled by the block's reciprocal absmax (one correctly rounded 1 / absmax per block, then a product per value;
        # bitsandbytes divides every value with the approximate __fdividef)
        inv1 = torch.where(a1 > 0, torch.ones_like(a1) / a1, torch.zeros(()))
        inv2 = torch.where(a2 > 0, torch.ones_like(a2) / a2, torch.zeros(()))
        x1, x2 = s1 * inv1[blk], s2 * inv2[blk]
        c1 = quantize_nearest(x1, self.qmap1).long()
        flip = torch.signbit(self.qmap1[c1]) != torch.signbit(s1)
        c1 = torch.where(flip, torch.where(s1 > 0, c1 + 1, c1 - 1), c1)
        c2 = quantize_nearest(x2, self.qmap2).long()
        self.state1[is8] = c1[is8].to(torch.uint8)
        self.state2[is8] = c2[is8].to(torch.uint8)
        self.absmax1, self.absmax2 = a1, a2
        return param.bfloat16()

    def moments(self):
        """(exp_avg, exp_avg_sq) de-quantised, for comparisons with a 32-bit optimizer"""
        mode, within = self._element_modes()
        blk = torch.arange(self.n) // BLOCK
        m = self.qmap1[self.state1.long()] * self.absmax1[blk]
        v = self.qmap2[self.state2.long()] * self.absmax2[blk]
        small = mode >= 0
        idx = (mode.long() + within)[small]
        m[small], v[small] = self.small_m[idx], self.small_v[idx]
        m[mode == MODE_SKIP] = 0
        v[mode == MODE_SKIP] = 0
        return m, v
