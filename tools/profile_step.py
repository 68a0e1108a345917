"""One eager UNet forward (CFG batch 2, 64x64 latent) for ncu: `ncu ... python tools/profile_step.py`.
Prints the number of b200sd launches of the warm-up so `-s` can skip them."""
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
x = torch.randn(2 * B, 4, 64, 64, device=dev)
ctx = torch.randn(2 * B, 77, 768, device=dev)
with torch.no_grad():
    for i in range(int(os.environ.get("ITERS", 2))):
        n0 = ops.launch_count()
        unet(x, 500 - i, ctx)
        torch.cuda.synchronize()
        print("launches this forward:", ops.launch_count() - n0, flush=True)
