"""Warm per-launch device time of every entry of the UNet launch plan (CFG batch 2, 64x64 latent), longest first:
    python tools/per_launch.py > profiles/r01_warm_per_launch_times.txt
Each entry is timed in its own CUDA graph of 4 back-to-back launches (Engine.profile), so the numbers carry no host gaps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd.unet import UNet2DConditionModel

B = int(os.environ.get("B", 1))
dev = torch.device("cuda:0")
torch.manual_seed(0)
unet = UNet2DConditionModel().to(dev).eval()
x = torch.randn(2 * B, 4, 64, 64, device=dev)
ctx = torch.randn(2 * B, 77, 768, device=dev)
with torch.no_grad():
    unet(x, 500, ctx)
    eng = next(iter(unet._engines.values()))
    acc, per_op, _total = eng.profile()
tot = sum(ms for _, ms, _ in per_op)
print(f"# {len(per_op)} launches, sum {tot:.3f} ms (isolated; the captured step is faster)")
for name, ms, fl in sorted(per_op, key=lambda t: -t[1]):
    print(f"{ms * 1e3:9.1f} us  {fl / ms / 1e9 if ms > 0 else 0:8.1f} TF/s  {name}")
