"""One training step (add_noise + UNet fwd + MSE + bwd + fused AdamW, batch B, 64x64 latent) for ncu:
`ncu ... python tools/profile_train_step.py`.  Prints the launches per step so `-s` can skip the warm-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
from b200sd.schedulers import DDPMScheduler
from b200sd.trainer import Trainer
from b200sd.unet import UNet2DConditionModel

B = int(os.environ.get("B", 8))
dev = torch.device("cuda:0")
torch.manual_seed(0)
unet = UNet2DConditionModel().to(dev)
tr = Trainer(unet, DDPMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000))
x0, noise = torch.randn(B, 4, 64, 64, device=dev), torch.randn(B, 4, 64, 64, device=dev)
t = torch.randint(0, 1000, (B,), device=dev)
ctx = torch.randn(B, 77, 768, device=dev)
for i in range(int(os.environ.get("ITERS", 2))):
    n0 = ops.launch_count()
    tr.train_step(x0, noise, t, ctx)
    torch.cuda.synchronize()
    print("launches this step:", ops.launch_count() - n0, flush=True)
