"""last GPU seconds of the round: load_state_dict of both flat optimizers after the validate-first rewrite (host logic, CUDA tensors)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd.train import FlatParams
from b200sd.trainer import FlatAdamW, FlatAdamW8bit
from b200sd.unet import UNet2DConditionModel
from oracle.unet_ref import TINY_OVERRIDES
torch.manual_seed(0)
m = UNet2DConditionModel(**TINY_OVERRIDES).to("cuda:0")
flat = FlatParams(m, torch.device("cuda:0"))
flat.attach_grads()
for cls in (FlatAdamW, FlatAdamW8bit):
    opt = cls(flat, lr=1e-3)
    flat.grad.normal_(0, 1e-2)
    opt.step()
    sd = opt.state_dict()
    o2 = cls(flat, lr=1.0)
    o2.load_state_dict(sd)
    a, b = opt.state[flat.master], o2.state[flat.master]
    assert o2.lr == 1e-3 and o2.steps == 1 and all(torch.equal(a[k], b[k]) and a[k].dtype == b[k].dtype for k in a if torch.is_tensor(a[k]))
    print(cls.__name__, "load_state_dict ok", {k: str(v.dtype) for k, v in b.items() if torch.is_tensor(v)})
try:
    FlatAdamW(flat).load_state_dict(sd)
except ValueError as e:
    print("cross-type load rejected:", e)
