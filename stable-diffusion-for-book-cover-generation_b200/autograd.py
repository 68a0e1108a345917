"""torch.autograd bridge of the training engine: `unet(noisy_latents, timesteps, encoder_hidden_states)`
inside a training loop (finetune_sd.py:480-494) records ONE autograd node whose backward replays the
backward launch plan of train.TrainEngine.

Two ways of delivering parameter gradients:
  * default ("autograd") -- backward() returns a gradient per parameter, so `param.grad`, gradient
    accumulation, `zero_grad()` and DistributedDataParallel hooks (accelerate wraps the UNet in DDP,
    finetune_sd.py:363, 386) behave exactly as with an eager module.  The gradients are views of the flat
    buffer; autograd clones them (one extra pass over 3.4 GB of fp32 at SD v1.5 size).
  * direct (`unet.enable_direct_gradients()`) -- `param.grad` ARE the views of the flat fp32 gradient
    buffer; the kernels accumulate into it and nothing is copied.  The buffer is zeroed when a backward
    finds a `param.grad` set to None (what `optimizer.zero_grad()` does) or by `unet.zero_grad()`.
    This is the mode the data-parallel trainer (trainer.py) uses: it allreduces the flat buffer itself.
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from .train import FlatParams, TrainEngine


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, sample, timestep, ehs, *params):
        ctx.engine = engine
        ctx.n_params = len(params)
        ctx.param_ids = [id(p) for p in params]
        ctx.need_ctx = ehs.requires_grad
        ctx.ehs_dtype = ehs.dtype
        out = engine.run_forward(sample.detach(), timestep, ehs.detach())
        # the engine keeps the activations the backward needs in per-geometry static buffers: a second forward of the
        # same geometry overwrites them, so each node remembers which forward it belongs to
        engine.forward_generation = getattr(engine, "forward_generation", 0) + 1
        ctx.generation = engine.forward_generation
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, d_out):
        eng = ctx.engine
        if ctx.generation != eng.forward_generation:
            raise RuntimeError("b200sd: backward() of a UNet forward whose saved activations were overwritten by a later forward "
                               "of the same geometry (the training engine keeps ONE set of static buffers per geometry): "
                               "call backward() before the next training forward")
        model, flat = eng.model, eng.flat
        direct = model._direct_grads
        if eng.train_weights:
            if not direct:
                flat.zero_grad()
            elif any(r.param.requires_grad and (r.param.grad is None or r.param.grad.data_ptr() != r.gview.data_ptr())
                     for r in flat.order):
                flat.zero_grad()
                flat.attach_grads()
        d_ctx = eng.run_backward(d_out.contiguous().float(), on_ready=model._grad_ready_hook if direct else None)
        if eng.train_weights and not direct:
            grads = tuple(flat.regs[i].gview if flat.regs[i].param.requires_grad else None for i in ctx.param_ids)
        else:
            grads = (None,) * ctx.n_params
        # d_ctx is a view of an engine buffer that the next backward overwrites: hand autograd its own copy
        return (None, None, None, d_ctx.to(ctx.ehs_dtype, copy=True) if ctx.need_ctx and d_ctx is not None else None) + grads


def ensure_flat(model, dev):
    """(Re)build the flat kernel-layout state when it does not exist, sits on another device, or the parameters were
    replaced (a .to() / load of new tensors un-homes them)."""
    if getattr(model, "_flat", None) is None or model._flat.device != dev or not model._flat.owns(model):
        model._flat = FlatParams(model, dev)
        model._train_engines = {}
    return model._flat.materialize()


def unet_forward_train(model, sample, timestep, ctx):
    if sample.requires_grad:
        raise NotImplementedError("b200sd: gradients w.r.t. the latent input are not implemented (the reference never needs them)")
    dev = sample.device
    ensure_flat(model, dev)
    N, _, H, W = sample.shape
    train_weights = any(p.requires_grad for p in model.parameters())
    key = (N, H, W, ctx.shape[1], dev.index, train_weights, bool(ctx.requires_grad))
    eng = model._train_engines.get(key)
    if eng is None:
        eng = TrainEngine(model, model._flat, N, H, W, ctx.shape[1], dev, train_weights=train_weights,
                          ctx_grad=bool(ctx.requires_grad))
        model._train_engines[key] = eng
    params = [p for p in model.parameters()] if train_weights else []
    out = _UNetFn.apply(eng, sample, timestep, ctx, *params)
    return out.to(sample.dtype) if out.dtype != sample.dtype else out
