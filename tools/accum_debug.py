import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd import ops
from b200sd.unet import UNet2DConditionModel
from oracle.unet_ref import TINY_OVERRIDES
torch.manual_seed(0)
DEV='cuda'
unet = UNet2DConditionModel(**TINY_OVERRIDES).to(DEV).train()
unet.enable_direct_gradients()
g = torch.Generator().manual_seed(5)
x = torch.randn(4, 4, 32, 32, generator=g).to(DEV); noise = torch.randn(4, 4, 32, 32, generator=g).to(DEV)
ctx = torch.randn(4, 77, 64, generator=g).to(DEV); t = torch.randint(0, 1000, (4,), generator=g).to(DEV)
grads = []
for i in range(3):
    ops._GN_RECOMPUTE = (i == 0)
    unet.zero_grad()
    ops.mse_loss(unet(x, t, ctx).sample, noise).backward()
    torch.cuda.synchronize()
    grads.append({n: p.grad.clone() for n, p in unet.named_parameters()})
order = [r.param for r in unet._flat.order]
names = {id(p): n for n, p in unet.named_parameters()}
print("flat order from the END (= backward completion order): recompute-vs-fast, fast-vs-fast")
for p in reversed(order):
    n = names[id(p)]
    a, b, c = grads[0][n], grads[1][n], grads[2][n]
    s = float(a.abs().max()) + 1e-30
    e1, e2 = float((a - b).abs().max()) / s, float((b - c).abs().max()) / s
    if e1 > 1e-4 or e2 > 1e-4:
        print(f"{e1:.3e} {e2:.3e} {n} {tuple(p.shape)}")
        break
    print(f"   ok {e1:.1e} {n}")
