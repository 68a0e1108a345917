"""The whole denoising step of StableDiffusionPipeline.__call__ as ONE CUDA graph (DDIM or PLMS + classifier-free guidance).

Reference call sites: the loop inside `pipeline(...)` -- inference.py:175-176, 342-351; finetune_sd.py:264-271 -- i.e. per
step `latent_model_input = cat([latents] * 2)`, `unet(...)`, `noise_pred_uncond + guidance_scale * (...)`, `scheduler.step`.

What the graph holds (nothing is launched eagerly between two steps, the host only calls cudaGraphLaunch):

    sampler_advance            in_t <- timesteps[cursor]; cursor += 1            (device-side step index)
      fork ──> lane 0: UNet plan over the UNCONDITIONAL half of the batch  ──┐
           └─> lane 1: UNet plan over the CONDITIONAL half                 ──┤  (lanes: independent launch chains on
      join <──────────────────────────────────────────────────────────────────┘   forked streams of the same graph)
    cfg_ddim_step_table        latents <- DDIM(eps_u + g (eps_c - eps_u)), coefficients = table[cursor], in place
    (PNDMScheduler / PLMS: cfg_plms_step_table -- weights, eps-ring slots and the saved-sample flags of this call = table[cursor];
     the 4-deep eps history is a device ring, PNDM's repeated first timestep reads the sample saved by the first call)

Lanes (`lanes` = 1, 2, or any even number that divides 2B into whole sub-batches of one half; default 1).  The two halves of
the CFG batch never meet before the combine, so they can run as two independent launch chains reading the SAME latent buffer
and the same weights.  Measured on B200 (profiles/r02b_lanes_sweep.txt): it LOSES at every batch -- 1 image 5.05 -> 5.67 ms
per step, 2 images 7.14 -> 7.58 (4 lanes 9.39), 4 images 11.15 -> 11.82, 8 images 20.14 -> 20.94 -- because the GEMM split
policy already spreads a half batch over the whole machine (smaller M -> more K slices), so the two chains contend for the
same SMs and each streams the weights on its own.  Kept as a tested option, off by default.

`host_step()` is the same graph with the host on both ends: H2D memcpy nodes for the latents and the text context out of
pinned host buffers, the context K/V projections (the context may change every call), the step, and a D2H memcpy node of
the new latents -- what bench.py reports as `e2e`.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import ops
from ._lib import B200SDError, check, lib
from .engine import Engine


def default_lanes(batch_images: int) -> int:
    """1 (measured: concurrent CFG halves lose at every batch size, see the module docstring); B200SD_LANES overrides."""
    env = os.environ.get("B200SD_LANES")
    return int(env) if env else 1


class CapturedSampler:
    def __init__(self, unet, scheduler, batch_images, h, w, S_ctx=77, guidance_scale=7.5, lanes=None):
        from .schedulers import DDIMScheduler, PNDMScheduler
        if not isinstance(scheduler, (DDIMScheduler, PNDMScheduler)):
            raise B200SDError("CapturedSampler covers DDIMScheduler and PNDMScheduler (PLMS)")
        self.plms = isinstance(scheduler, PNDMScheduler)
        if scheduler.num_inference_steps is None:
            raise ValueError("call scheduler.set_timesteps(n) first")
        if unet._precision != "bf16":
            raise B200SDError("CapturedSampler runs the bf16 plan")
        B = int(batch_images)
        L = default_lanes(B) if lanes is None else int(lanes)
        if L < 1 or (L > 1 and (L % 2 or B % (L // 2))):
            raise ValueError(f"lanes={L} does not split the {2 * B}-sample CFG batch into whole sub-batches of one half")
        self.unet, self.scheduler, self.B, self.h, self.w, self.S, self.lanes = unet, scheduler, B, h, w, S_ctx, L
        self.guidance = float(guidance_scale)
        dev = self.device = unet.device
        unet._ensure_packed()
        self._pack_gen = unet._pack_gen
        cfg = unet.config
        f32 = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            self.latents = torch.zeros(B, cfg.in_channels, h, w, **f32)        # updated in place by every step
            self.eps = torch.zeros(2 * B, cfg.out_channels, h, w, **f32)       # [uncond | cond] noise predictions
            self.in_t = torch.zeros(2 * B, **f32)
            self.ctx = torch.zeros(2 * B, S_ctx, cfg.cross_attention_dim, **f32)
            ts = [float(t) for t in scheduler.timesteps.tolist()]
            self.n_steps = len(ts)
            self.t_table = torch.tensor(ts, **f32)
            if self.plms:
                self.coef_table = torch.tensor(scheduler.plms_plan(), **f32).contiguous()       # [calls][12]
                self.saved = torch.zeros_like(self.latents)                                        # the first call's sample
                self.ring = torch.zeros(4, self.latents.numel(), **f32)                            # eps history
            else:
                self.coef_table = torch.tensor([scheduler._coefs(int(t)) for t in ts], **f32).contiguous()
            self.cursor = torch.zeros(2, dtype=torch.int32, device=dev)
            self.streams = [torch.cuda.Stream(device=dev) for _ in range(L)]
            # lane l of L > 1: half = l // (L/2) (0 = unconditional), images [j*c, (j+1)*c) of that half
            self.engines, self._slices = [], []
            if L == 1:
                self.x2 = torch.zeros(2 * B, cfg.in_channels, h, w, **f32)
                with torch.cuda.stream(self.streams[0]):
                    self.engines.append(Engine(unet, 2 * B, h, w, S_ctx, dev, io=(self.x2, self.in_t, self.eps)))
                self._slices.append((0, 2 * B))
            else:
                per, c = L // 2, B // (L // 2)
                for l in range(L):
                    half, j = divmod(l, per)
                    lo = half * B + j * c
                    io = (self.latents[j * c:(j + 1) * c], self.in_t[lo:lo + c], self.eps[lo:lo + c])
                    # built under the lane's own stream: the split-K workspace of the GEMMs is per (device, stream)
                    with torch.cuda.stream(self.streams[l]):
                        self.engines.append(Engine(unet, c, h, w, S_ctx, dev, io=io))
                    self._slices.append((lo, lo + c))
            self.graph = None
            self.host_graph = None
            self.kernels_per_step = 0
            self._ctx_key = None

    # -- building blocks ----------------------------------------------------------------------------
    def _fork_join(self, fn):
        """Run fn(lane) for every lane: lane 0 on the current stream, the others on their own streams forked from / joined
        into it with events (inside a capture these become the graph's parallel branches)."""
        cur = torch.cuda.current_stream()
        if self.lanes == 1:
            fn(0)
            return
        ev = torch.cuda.Event()
        ev.record(cur)
        done = []
        for l in range(1, self.lanes):
            st = self.streams[l]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                fn(l)
                e = torch.cuda.Event()
                e.record(st)
            done.append(e)
        fn(0)
        for e in done:
            cur.wait_event(e)

    def _project_context(self):
        def lane(l):
            eng = self.engines[l]
            lo, hi = self._slices[l]
            eng.in_ctx.copy_(self.ctx[lo:hi].reshape(-1, self.ctx.shape[-1]))
            for op in eng.ctx_plan:
                op()
        self._fork_join(lane)

    def _step_ops(self):
        check(lib().b200sd_sampler_advance(self.t_table.data_ptr(), self.n_steps, self.cursor.data_ptr(), self.in_t.data_ptr(),
                                           self.in_t.numel(), ops._stream()), "sampler_advance")
        if self.lanes == 1:
            self.x2[:self.B].copy_(self.latents)
            self.x2[self.B:].copy_(self.latents)
        self._fork_join(lambda l: self.engines[l]._run_plan())
        B = self.B
        if self.plms:
            check(lib().b200sd_cfg_plms_step_table(self.eps[:B].data_ptr(), self.eps[B:].data_ptr(), self.latents.data_ptr(),
                                                   self.saved.data_ptr(), self.ring.data_ptr(), self.latents.numel(), self.guidance,
                                                   self.coef_table.data_ptr(), self.cursor.data_ptr(), ops._stream()),
                  "cfg_plms_step_table")
            return
        check(lib().b200sd_cfg_ddim_step_table(self.eps[:B].data_ptr(), self.eps[B:].data_ptr(), self.latents.data_ptr(),
                                               self.latents.data_ptr(), None, self.latents.numel(), self.guidance,
                                               self.coef_table.data_ptr(), self.cursor.data_ptr(), ops.F32, ops.F32,
                                               ops._stream()), "cfg_ddim_step_table")

    def _capture(self, body):
        """Warm up `body` on the capture stream (lazy one-time setup: function attributes, per-stream workspaces), then capture."""
        s0 = self.streams[0]
        s0.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s0):
            body()
        s0.synchronize()
        cur = int(self.cursor[0])      # the warm-up advanced the device-side step index: put it back
        self.cursor[0] = (cur - 1) % self.n_steps
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        before = ops.launch_count()
        with torch.cuda.graph(g, stream=s0):
            body()
        return g, ops.launch_count() - before

    # -- public surface -----------------------------------------------------------------------------
    def reset(self, step=0):
        """the next step() executes scheduler.timesteps[step]"""
        self.cursor.copy_(torch.tensor([step % self.n_steps, 0], dtype=torch.int32))

    def set_context(self, ctx2):
        """ctx2 = cat([uncond, cond]) (2B, S, 768): projected once to every cross-attention's K/V"""
        if tuple(ctx2.shape) != tuple(self.ctx.shape):
            raise ValueError(f"encoder_hidden_states must be {tuple(self.ctx.shape)}, got {tuple(ctx2.shape)}")
        with torch.cuda.device(self.device):
            self.ctx.copy_(ctx2)
            self._project_context()

    def set_latents(self, latents):
        self.latents.copy_(latents)

    def step(self):
        """one denoising iteration: ONE cudaGraphLaunch, no eager kernels"""
        with torch.cuda.device(self.device):
            if self.graph is None:
                keep = self.latents.clone()
                self.graph, self.kernels_per_step = self._capture(self._step_ops)
                self.latents.copy_(keep)       # the warm-up run updated the latents in place
            self.graph.replay()

    @torch.no_grad()
    def run(self, latents, ctx2):
        """latents (B,4,h,w) ~ N(0,1) -> denoised latents after scheduler.num_inference_steps steps"""
        if self.unet._stale(self._pack_gen):
            raise B200SDError("the UNet's weights changed after this sampler was built: build a new CapturedSampler")
        self.set_context(ctx2.float())
        self.set_latents(latents.float() * self.scheduler.init_noise_sigma)
        self.reset(0)
        for _ in range(self.n_steps):
            self.step()
        return self.latents.clone()

    def bind_host(self, lat_host, ctx_host, out_host):
        """Capture the host-facing step over these PINNED host buffers (latents in, context in, new latents out)."""
        for t in (lat_host, ctx_host, out_host):
            if not t.is_pinned():
                raise B200SDError("host_step needs pinned host buffers")

        def body():
            self.latents.copy_(lat_host, non_blocking=True)
            self.ctx.copy_(ctx_host, non_blocking=True)
            self._project_context()
            self._step_ops()
            out_host.copy_(self.latents, non_blocking=True)
        with torch.cuda.device(self.device):
            self.host_graph, self.kernels_per_host_step = self._capture(body)
        self._host = (lat_host, ctx_host, out_host)

    def host_step(self):
        """H2D(latents, context) -> context K/V -> step -> D2H(latents): one graph launch + one stream synchronize"""
        with torch.cuda.device(self.device):
            self.host_graph.replay()
            torch.cuda.current_stream().synchronize()
