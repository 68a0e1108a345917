"""Eager UNet forwards (CFG batch 2, 64x64 latent) for ncu: `ncu ... python tools/profile_step.py [names.txt]`.
Writes one line per KERNEL LAUNCH of the last forward (plan order: "<kind>\t<name>") to names.txt, so that an ncu launch list
can be joined with the plan entries by position (tools/ncu_join.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
from b200sd.unet import UNet2DConditionModel

B = int(os.environ.get("B", 1))
dev = torch.device("cuda:0")
torch.manual_seed(0)
unet = UNet2DConditionModel().to(dev).eval()
unet.use_cuda_graph = False
h, w = (96, 64) if os.environ.get("PORTRAIT") else (64, 64)
x = torch.randn(2 * B, 4, h, w, device=dev)
ctx = torch.randn(2 * B, 77, 768, device=dev)
with torch.no_grad():
    for i in range(int(os.environ.get("ITERS", 2))):
        n0 = ops.launch_count()
        unet(x, 500 - i, ctx)
        torch.cuda.synchronize()
        print("launches this forward:", ops.launch_count() - n0, flush=True)
    if len(sys.argv) > 1:
        # one more forward, op by op, counting the launches of every plan entry
        eng = next(iter(unet._engines.values()))
        lines = []
        for op, (kind, flops, name) in zip(eng.plan, eng.plan.meta):
            if kind == "tap":
                continue
            n0 = ops.launch_count()
            op()
            for _ in range(ops.launch_count() - n0):
                lines.append(f"{kind}\t{name or kind}\t{flops:.0f}")
        torch.cuda.synchronize()
        with open(sys.argv[1], "w") as f:
            f.write("\n".join(lines) + "\n")
        print("launches named:", len(lines), flush=True)
