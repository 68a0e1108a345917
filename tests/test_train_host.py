"""CPU tests of the training path's host logic: the flat kernel-layout parameter state (re-homing keeps the
diffusers state-dict surface intact) and the bucketed gradient allreduce over a 2-rank gloo group
(the N>1 path of BASELINE config 3; NCCL on the GPU box, gloo here)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _tiny():
    from b200sd.unet import UNet2DConditionModel
    from oracle.unet_ref import TINY_OVERRIDES
    torch.manual_seed(0)
    return UNet2DConditionModel(**TINY_OVERRIDES)


def test_flat_params_rehoming_keeps_the_module_surface():
    from b200sd.train import FlatParams
    m = _tiny()
    before = {k: v.clone() for k, v in m.state_dict().items()}
    n_params = sum(p.numel() for p in m.parameters())
    flat = FlatParams(m, torch.device("cpu"))
    assert flat.total >= n_params and flat.total % 64 == 0
    assert flat.owns(m)
    after = m.state_dict()
    assert list(after) == list(before)
    for k in before:
        assert after[k].shape == before[k].shape and torch.equal(after[k], before[k]), k
    # every parameter is a view of the ONE master buffer; 3x3 conv weights are stored [Cout][ky][kx][Cin]
    base = flat.master.untyped_storage().data_ptr()
    assert all(p.data.untyped_storage().data_ptr() == base for p in m.parameters())
    w = m.down_blocks[0].resnets[0].conv1.weight
    r = flat.reg(w)
    assert torch.equal(flat.master[r.off:r.off + r.numel].view(w.shape[0], 3, 3, w.shape[1]), w.detach().permute(0, 2, 3, 1))
    # fused regions are adjacent
    a = m.down_blocks[0].attentions[0].transformer_blocks[0].attn1
    qkv = flat.span([a.to_q.weight, a.to_k.weight, a.to_v.weight], "master")
    assert torch.equal(qkv, torch.cat([a.to_q.weight, a.to_k.weight, a.to_v.weight]).detach())
    tp = flat.span([r_.time_emb_proj.bias for _, r_ in m._iter_resnets()], "master")
    assert torch.equal(tp, torch.cat([r_.time_emb_proj.bias for _, r_ in m._iter_resnets()]).detach())
    # gradients: views of the flat buffer with the parameter's own strides; an in-place optimizer update
    # of the re-homed parameter lands in the master buffer
    flat.attach_grads()
    assert w.grad.shape == w.shape and w.grad.stride() == w.stride()
    flat.grad.fill_(1.0)
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    opt.step()
    assert torch.allclose(flat.master[r.off:r.off + r.numel], (before["down_blocks.0.resnets.0.conv1.weight"] - 0.5)
                          .permute(0, 2, 3, 1).reshape(-1))
    # load_state_dict writes through the views
    m.load_state_dict(before)
    assert flat.owns(m) and torch.equal(m.state_dict()["conv_out.weight"], before["conv_out.weight"])
    # layout order = forward execution order (the backward completes the buffer back to front)
    offs = [flat.reg(p).off for p in (m.time_embedding.linear_1.weight, m.conv_in.weight, m.down_blocks[0].resnets[0].norm1.weight,
                                      m.down_blocks[0].attentions[0].norm.weight, m.down_blocks[0].resnets[1].norm1.weight,
                                      m.mid_block.resnets[0].norm1.weight, m.up_blocks[3].attentions[2].proj_out.bias,
                                      m.conv_norm_out.weight, m.conv_out.bias)]
    assert offs == sorted(offs) and offs[0] == 0


def test_bucket_reducer_single_process_ranges():
    from b200sd.trainer import BucketReducer
    flat = torch.ones(1000)
    r = BucketReducer(flat, bucket_bytes=1000)      # 250 floats per bucket
    for off in (900, 800, 650, 640, 300, 0):
        r.on_ready(off)
    r.finish()
    assert r.ranges == [(650, 1000), (300, 650), (0, 300)]
    r.begin()
    r.on_ready(990)
    r.finish()                                       # the tail is flushed by finish()
    assert r.ranges == [(0, 1000)]


def test_clip_flat_rehoming_keeps_the_transformers_surface():
    """ClipFlat: every CLIP parameter is a view of one master buffer, q|k|v weights and biases adjacent (one fused QKV GEMM),
    state_dict unchanged."""
    from b200sd.clip import CLIPTextModel, ClipFlat
    torch.manual_seed(0)
    m = CLIPTextModel(vocab_size=300, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=1)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    flat = ClipFlat(m, torch.device("cpu"))
    assert flat.owns(m) and flat.total % 64 == 0
    after = m.state_dict()
    assert list(after) == list(before) and all(torch.equal(after[k], before[k]) for k in before)
    a = m.text_model.encoder.layers[1].self_attn
    assert torch.equal(flat.span([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], "master"),
                       torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight]).detach())
    assert torch.equal(flat.span([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], "master"),
                       torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias]).detach())
    tm = m.text_model
    offs = [flat.reg(p).off for p in (tm.embeddings.token_embedding.weight, tm.encoder.layers[0].layer_norm1.weight,
                                      tm.encoder.layers[1].mlp.fc2.bias, tm.final_layer_norm.bias)]
    assert offs == sorted(offs) and offs[0] == 0
    flat.attach_grads()
    assert a.q_proj.weight.grad.shape == a.q_proj.weight.shape


def test_allreduce_in_chunks_ranges_single_process():
    from b200sd.trainer import allreduce_in_chunks
    assert allreduce_in_chunks(torch.ones(10), 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert allreduce_in_chunks(torch.ones(8), 1) == [(0, 8)]
    assert allreduce_in_chunks(torch.ones(3), 8) == [(0, 1), (1, 2), (2, 3)]


def _chunk_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from b200sd.trainer import allreduce_in_chunks
        flat = torch.arange(1001, dtype=torch.float32) * (rank + 1)
        ranges = allreduce_in_chunks(flat, 4)
        want = torch.arange(1001, dtype=torch.float32) * sum(range(1, world + 1))
        q.put((rank, bool(torch.equal(flat, want)), ranges))
    finally:
        dist.destroy_process_group()


def test_allreduce_in_chunks_two_ranks_gloo():
    """the text-encoder trainer's collective (BASELINE config 4 at N > 1), gloo here / NCCL on the GPU box"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_chunk_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ranges in res:
        assert ok, f"rank {rank}: chunked allreduce result wrong"
        assert ranges == [(0, 251), (251, 502), (502, 753), (753, 1001)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from b200sd.trainer import BucketReducer
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        red = BucketReducer(flat, bucket_bytes=1200)
        for off in (700, 500, 200, 0):
            red.on_ready(off)
        red.finish()
        want = torch.arange(1000, dtype=torch.float32) * sum(range(1, world + 1))
        ok = bool(torch.equal(flat, want))
        # finish(wait=False) + pending(): what Trainer uses to interleave the optimizer with the collectives
        flat2 = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        red2 = BucketReducer(flat2, bucket_bytes=1200)
        red2.on_ready(600)
        red2.finish(wait=False)
        pend = red2.pending()
        for a, b, work in pend:
            work.wait()
            ok = ok and bool(torch.equal(flat2[a:b], want[a:b]))
        ok = ok and [(a, b) for a, b, _ in pend] == [(600, 1000), (0, 600)] and red2.works == []
        q.put((rank, ok, red.ranges))
    finally:
        dist.destroy_process_group()


def test_bucket_reducer_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ranges in res:
        assert ok, f"rank {rank}: allreduce result wrong"
        assert ranges == [(700, 1000), (200, 700), (0, 200)]


def _torch_adamw_stand_in(monkeypatch):
    """ops.adamw_step is a CUDA kernel; on the CPU the same update (torch.optim.AdamW's formulas, decoupled decay, grads zeroed)
    stands in so that the optimizer-surface plumbing of FlatAdamW can be exercised without a GPU."""
    from b200sd import ops

    def adamw_step(param, grad, exp_avg, exp_avg_sq, weights_bf16, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0,
                   zero_grad=True):
        g = grad * grad_scale
        param.mul_(1 - lr * weight_decay)
        exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
        exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (exp_avg_sq / (1 - beta2 ** step)).sqrt_().add_(eps)
        param.addcdiv_(exp_avg, denom, value=-lr / (1 - beta1 ** step))
        if weights_bf16 is not None:
            weights_bf16.copy_(param)
        if zero_grad:
            grad.zero_()
    monkeypatch.setattr(ops, "adamw_step", adamw_step)


def test_flat_adamw_is_a_torch_optimizer_driven_by_the_reference_lr_schedule(monkeypatch):
    """finetune_sd.py:421-422, 577: CosineAnnealingLR(optimizer, T_max, eta_min=1e-6) + scheduler.step() after optimizer.step().
    FlatAdamW takes the lr from param_groups[0] at every step, so torch's own scheduler drives it; the parameters follow
    torch.optim.AdamW under the same schedule; state_dict round-trips; rebind() keeps moments, step count AND the scheduler."""
    from b200sd.train import FlatParams
    from b200sd.trainer import FlatAdamW
    _torch_adamw_stand_in(monkeypatch)
    m = _tiny()
    m.mark_weights_changed = lambda: None
    ref = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    flat = FlatParams(m, torch.device("cpu"))
    flat.attach_grads()
    opt = FlatAdamW(flat, lr=1e-3, betas=(0.9, 0.99), weight_decay=0.1)
    assert isinstance(opt, torch.optim.Optimizer) and opt.param_groups[0]["params"][0] is flat.master
    ropt = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.99), weight_decay=0.1)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=4, eta_min=1e-6)
    rsched = torch.optim.lr_scheduler.CosineAnnealingLR(ropt, T_max=4, eta_min=1e-6)
    gen = torch.Generator().manual_seed(0)
    lrs = []
    for it in range(3):
        for p, q in zip(m.parameters(), ref):
            g = torch.randn(p.shape, generator=gen)
            p.grad.copy_(g)
            q.grad = g.clone()
        if it == 1:      # the trainers' bucket-by-bucket form of the same step
            opt.begin_step()
            half = flat.total // 2 // 64 * 64
            opt.step_range(half, flat.total)
            opt.step_range(0, half)
            opt.end_step()
        else:
            opt.step()
        ropt.step()
        sched.step()
        rsched.step()
        lrs.append(opt.lr)
        assert opt.lr == ropt.param_groups[0]["lr"] and opt.lr < 1e-3
        assert float(flat.grad.abs().max()) == 0.0
    assert lrs == sorted(lrs, reverse=True) and opt.steps == 3
    for p, q in zip(m.parameters(), ref):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-5, atol=1e-7)
    # state_dict round trip into a fresh optimizer over the same buffers
    sd = opt.state_dict()
    assert sd["state"][0]["step"] == 3 and sd["param_groups"][0]["lr"] == opt.lr
    opt2 = FlatAdamW(flat, lr=5.0)
    opt2.load_state_dict(sd)
    assert opt2.lr == opt.lr and opt2.steps == 3 and torch.equal(opt2.exp_avg, opt.exp_avg) and opt2.betas == (0.9, 0.99)
    # the flat buffers are rebuilt (unet.to() un-homes the parameters): same object, same state, scheduler still attached
    exp_avg = opt.exp_avg.clone()
    flat_b = FlatParams(m, torch.device("cpu"))
    assert opt.rebind(flat_b) is opt and opt.param_groups[0]["params"][0] is flat_b.master
    assert opt.steps == 3 and torch.equal(opt.exp_avg, exp_avg) and len(opt.state) == 1
    sched.step()
    assert opt.lr == sched.get_last_lr()[0] < lrs[-1]
    with pytest.raises(ValueError):
        FlatAdamW(flat_b, lr=-1.0)


def _oracle_adamw8bit_stand_in(monkeypatch):
    """ops.adamw8bit_step is a CUDA kernel; on the CPU the oracle's step (the kernel is bit-exact against it on the GPU:
    tests/test_optim8bit_gpu.py) stands in behind the same argument list, so FlatAdamW8bit's host logic -- chunk table, state
    buffers, range-wise stepping, state_dict, rebind -- runs without a GPU."""
    from b200sd import ops
    from oracle import adam8bit_ref as A
    calls = []

    def adamw8bit_step(param, grad, state1, state2, absmax1, absmax2, qmap1, qmap2, chunk_mode, small_m, small_v, wb, lr, beta1, beta2,
                       eps, weight_decay, step, grad_scale=1.0, zero_grad=False):
        ref = A.AdamW8bitRef(param.numel(), chunk_mode, lr=lr, betas=(beta1, beta2), eps=eps, weight_decay=weight_decay)
        assert torch.equal(ref.qmap1, qmap1) and torch.equal(ref.qmap2, qmap2)
        ref.state1, ref.state2, ref.absmax1, ref.absmax2 = state1, state2, absmax1.clone(), absmax2.clone()
        n_small = ref.small_m.numel()
        ref.small_m, ref.small_v, ref.steps = small_m[:max(n_small, 0)], small_v[:max(n_small, 0)], step - 1
        wb.copy_(torch.where((ref.chunk_mode != A.MODE_SKIP).repeat_interleave(64), ref.step(param, grad, grad_scale, zero_grad), wb))
        absmax1.copy_(ref.absmax1)
        absmax2.copy_(ref.absmax2)
        calls.append(step)
    monkeypatch.setattr(ops, "adamw8bit_step", adamw8bit_step)
    return calls


def test_flat_adamw8bit_host_logic(monkeypatch):
    """Trainer(optim_bits=8) == bnb.optim.AdamW8bit(..., min_8bit_size=16384) (finetune_sd.py:300, 407-410): tensors below
    16 384 elements keep fp32 moments at disjoint offsets of a compact buffer, frozen tensors are skipped, everything else is
    8-bit; ~2 B of state per parameter; one launch per step, also in the trainers' range-by-range form; lr schedule, state_dict
    (u8 codes stay u8) and rebind like FlatAdamW."""
    from b200sd.train import FlatParams
    from b200sd.trainer import FlatAdamW, FlatAdamW8bit, _make_optimizer
    calls = _oracle_adamw8bit_stand_in(monkeypatch)
    m = _tiny()
    m.mark_weights_changed = lambda: None
    m.conv_in.weight.requires_grad_(False)
    flat = FlatParams(m, torch.device("cpu"))
    flat.attach_grads()
    opt = _make_optimizer(flat, 8, dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2), 16384)
    assert isinstance(opt, FlatAdamW8bit) and isinstance(opt, FlatAdamW) and isinstance(opt, torch.optim.Optimizer)
    assert type(_make_optimizer(flat, 32, dict(lr=1e-3), 16384)) is FlatAdamW
    with pytest.raises(ValueError):
        _make_optimizer(flat, 16, {}, 16384)
    # the chunk table
    mode = opt.chunk_mode
    assert mode.numel() == flat.total // 64 and mode.dtype == torch.int32
    small_offsets = []
    for r in flat.order:
        chunks = mode[r.off // 64:(r.off + r.numel + 63) // 64]
        if not r.param.requires_grad:
            assert bool((chunks == -2).all())
        elif r.numel < 16384:
            assert bool((chunks >= 0).all()) and bool((chunks[1:] - chunks[:-1] == 64).all())
            small_offsets += chunks.tolist()
        else:
            assert bool((chunks == -1).all())
    assert len(set(small_offsets)) == len(small_offsets) and sorted(small_offsets) == list(range(0, 64 * len(small_offsets), 64))
    st = opt.state[flat.master]
    assert st["exp_avg"].numel() == 64 * len(small_offsets) and st["state1"].dtype == torch.uint8
    assert st["absmax1"].numel() == (flat.total + 2047) // 2048
    assert opt.state_bytes() < 2.5 * flat.total
    # steps: one launch each, frozen parameter untouched, lr schedule applies
    frozen = m.conv_in.weight.detach().clone()
    w = m.down_blocks[0].resnets[0].conv1.weight
    before = w.detach().clone()
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=4, eta_min=1e-6)
    gen = torch.Generator().manual_seed(0)
    for it in range(3):
        flat.grad.copy_(torch.randn(flat.total, generator=gen) * 1e-2)
        if it == 1:
            opt.begin_step()
            opt.step_range(flat.total // 2 // 64 * 64, flat.total, grad_scale=0.5)
            opt.step_range(0, flat.total // 2 // 64 * 64, grad_scale=0.5)
            with pytest.raises(ValueError):
                opt.step_range(0, 64, grad_scale=0.25)
            opt.end_step()
        else:
            opt.step(grad_scale=0.5)
        sched.step()
    assert calls == [1, 2, 3] and opt.steps == 3 and opt.lr < 1e-3
    assert torch.equal(m.conv_in.weight.detach(), frozen) and not torch.equal(w.detach(), before)
    live = (mode != -2).repeat_interleave(64)
    assert float(flat.grad[live].abs().max()) == 0.0 and int(st["state1"].max()) > 0 and float(st["exp_avg"].abs().max()) > 0
    assert torch.equal(flat.wb[live], flat.master[live].bfloat16())
    # state_dict round trip (torch casts floating state to the parameter's dtype: the codes must come back as u8, bit for bit)
    sd = opt.state_dict()
    opt2 = FlatAdamW8bit(flat, lr=1.0)
    opt2.load_state_dict(sd)
    st2 = opt2.state[flat.master]
    assert opt2.steps == 3 and opt2.lr == opt.lr and st2["state1"].dtype == torch.uint8
    assert all(torch.equal(st2[k], st[k]) for k in ("state1", "state2", "absmax1", "absmax2", "exp_avg", "exp_avg_sq"))
    other = FlatAdamW(flat, lr=1e-3)
    with pytest.raises(ValueError):
        other.load_state_dict(sd)                                # an 8-bit state does not load into the fp32-moment optimizer ...
    assert other.lr == 1e-3 and other.steps == 0 and other.exp_avg.numel() == flat.total      # ... which stays as it was
    # rebuild of the flat buffers: same object, same state
    codes = st["state1"].clone()
    flat_b = FlatParams(m, torch.device("cpu"))
    assert opt.rebind(flat_b) is opt and torch.equal(opt.state[flat_b.master]["state1"], codes) and opt.steps == 3
