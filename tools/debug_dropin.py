"""debug: which of {DDP wrap, autocast, fp16 inputs} breaks the gradient; traceback of the frozen-fp16 case"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_dropin_gpu as T
DEV = "cuda:0"

def grads_cos(unet, oracle):
    named = dict(unet.named_parameters())
    worst = []
    for n, p in oracle.named_parameters():
        g = named[n].grad
        if g is None:
            worst.append((-1.0, n)); continue
        c = float(F.cosine_similarity(g.flatten().double().cpu(), p.grad.flatten().double(), dim=0))
        worst.append((c, n))
    worst.sort()
    ref = torch.cat([p.grad.flatten() for _, p in oracle.named_parameters()])
    got = torch.cat([named[n].grad.flatten().cpu() for n, _ in oracle.named_parameters()])
    return float(F.cosine_similarity(got.double(), ref.double(), dim=0)), worst[:5]

def run(ddp, autocast, half, warm=False, bucket_view=False):
    oracle, unet = T._pair(train=True)
    x, noise, ctx, t = T._inputs(2)
    oracle.zero_grad(set_to_none=True)
    F.mse_loss(oracle(x, t, ctx).sample, noise).backward()
    if warm:
        from b200sd import ops
        ops.mse_loss(unet(x.to(DEV), t.to(DEV), ctx.to(DEV)).sample, noise.to(DEV)).backward()
        unet.zero_grad(set_to_none=True)
    m = torch.nn.parallel.DistributedDataParallel(unet, device_ids=[0], gradient_as_bucket_view=bucket_view) if ddp else unet
    xx, cc = x.to(DEV), ctx.to(DEV)
    if half: xx, cc = xx.half(), cc.half()
    import contextlib
    cm = torch.autocast("cuda", dtype=torch.float16) if autocast else contextlib.nullcontext()
    with cm:
        pred = m(xx, t.to(DEV), cc).sample
        loss = F.mse_loss(pred.float(), noise.to(DEV), reduction="none").mean([1, 2, 3]).mean()
    loss.backward()
    print(f"ddp={ddp} autocast={autocast} half={half} warm={warm} bucket_view={bucket_view}: loss {float(loss):.5f}", grads_cos(unet, oracle), flush=True)

os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29578")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
for cfg in [(True, False, False), (True, False, False, True), (True, False, False, False, True), (True, False, False, True, True)]:
    try:
        run(*cfg)
    except Exception:
        traceback.print_exc()
# frozen fp16 UNet -> context gradient: which ingredient breaks it
def frozen(half_weights, half_inputs):
    from b200sd import ops
    oracle, unet = T._pair()
    x, noise, ctx, t = T._inputs(2)
    c_ref = ctx.clone().requires_grad_(True)
    F.mse_loss(oracle(x, t, c_ref).sample, noise).backward()
    unet = unet.requires_grad_(False)
    if half_weights:
        unet = unet.to(DEV, dtype=torch.float16)
    c = ctx.to(DEV)
    xx, nn_ = x.to(DEV), noise.to(DEV)
    if half_inputs:
        c, xx, nn_ = c.half(), xx.half(), nn_.half()
    c.requires_grad_(True)
    pred = unet(xx, t.to(DEV), c).sample
    with torch.no_grad():
        want = oracle(x, t, ctx).sample
    print("  fwd rel", float((pred.detach().float().cpu() - want).abs().max() / want.abs().max()), pred.dtype)
    ops.mse_loss(pred, nn_).backward()
    cos = float(F.cosine_similarity(c.grad.float().flatten().cpu(), c_ref.grad.flatten(), dim=0))
    print(f"frozen half_weights={half_weights} half_inputs={half_inputs}: ctx-grad cos {cos:.5f} |g| {float(c.grad.float().abs().max()):.3e} ref {float(c_ref.grad.abs().max()):.3e}", flush=True)
for cfg in [(False, False), (True, False), (False, True), (True, True)]:
    try:
        frozen(*cfg)
    except Exception:
        traceback.print_exc()
dist.destroy_process_group()
