"""Debug helper (not a test): per-block relative error of the CUDA UNet against the fp32 oracle run on
the same GPU.  `python tests/debug_unet_layers.py`"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sd.unet import UNet2DConditionModel
from oracle.unet_ref import make_oracle_unet

DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
oracle = make_oracle_unet(0, sharpen_attention=float(os.environ.get("SHARPEN", 4.0)))
ours = UNet2DConditionModel(); ours.load_state_dict(oracle.state_dict()); ours = ours.to(DEV).eval()
oracle = oracle.to(DEV)
g = torch.Generator().manual_seed(0)
N = int(os.environ.get("N", 2))
x = torch.randn(N, 4, 64, 64, generator=g).to(DEV); ctx = torch.randn(N, 77, 768, generator=g).to(DEV)
ref = {}
def hook(name):
    def f(mod, inp, out):
        ref[name] = (out[0] if isinstance(out, tuple) else out).detach().float()
    return f
oracle.conv_in.register_forward_hook(hook("conv_in"))
for i, b in enumerate(oracle.down_blocks):
    for j, r in enumerate(b.resnets): r.register_forward_hook(hook(f"down{i}.res{j}"))
    if b.attentions is not None:
        for j, a in enumerate(b.attentions): a.register_forward_hook(hook(f"down{i}.attn{j}"))
    if b.downsamplers is not None: b.downsamplers[0].register_forward_hook(hook(f"down{i}.ds"))
for j, r in enumerate(oracle.mid_block.resnets): r.register_forward_hook(hook(f"mid.res{j}"))
oracle.mid_block.attentions[0].register_forward_hook(hook("mid.attn0"))
for i, b in enumerate(oracle.up_blocks):
    for j, r in enumerate(b.resnets): r.register_forward_hook(hook(f"up{i}.res{j}"))
    if b.attentions is not None:
        for j, a in enumerate(b.attentions): a.register_forward_hook(hook(f"up{i}.attn{j}"))
    if b.upsamplers is not None: b.upsamplers[0].register_forward_hook(hook(f"up{i}.us"))
with torch.no_grad():
    want = oracle(x, 500, ctx).sample
    ours.use_cuda_graph = False
    _ = ours(x, 500, ctx).sample           # builds the engine
    eng = next(iter(ours._engines.values()))
    eng.debug = True
    got = ours(x, 500, ctx).sample
for name, t in eng.debug_out.items():
    r = ref[name]
    err = float((t - r).abs().max() / r.abs().max())
    print(f"{name:14s} shape {tuple(t.shape)}  max-rel {err:.5f}   ref absmax {float(r.abs().max()):.3f} std {float(r.std()):.3f}")
print("final max-rel", float((got - want).abs().max() / want.abs().max()))

# calibration: the same oracle module in torch eager bf16 (cuDNN / cuBLAS) vs fp32
import copy
ref32 = dict(ref)
ob = copy.deepcopy(oracle).bfloat16()
ref.clear()
def rehook(mod_o, mod_b):
    pass
hooks = {}
def hook2(name):
    def f(mod, inp, out):
        hooks[name] = (out[0] if isinstance(out, tuple) else out).detach().float()
    return f
ob.conv_in.register_forward_hook(hook2("conv_in"))
for i, b in enumerate(ob.down_blocks):
    for j, r in enumerate(b.resnets): r.register_forward_hook(hook2(f"down{i}.res{j}"))
    if b.attentions is not None:
        for j, a in enumerate(b.attentions): a.register_forward_hook(hook2(f"down{i}.attn{j}"))
for j, r in enumerate(ob.mid_block.resnets): r.register_forward_hook(hook2(f"mid.res{j}"))
for i, b in enumerate(ob.up_blocks):
    for j, r in enumerate(b.resnets): r.register_forward_hook(hook2(f"up{i}.res{j}"))
    if b.attentions is not None:
        for j, a in enumerate(b.attentions): a.register_forward_hook(hook2(f"up{i}.attn{j}"))
with torch.no_grad():
    wb = ob(x.bfloat16(), 500, ctx.bfloat16()).sample.float()
for name in ["conv_in", "down0.attn1", "down1.attn1", "down2.attn1", "mid.res1", "up1.attn2", "up2.attn2", "up3.attn2"]:
    r = ref32[name]; t = hooks[name]
    print(f"torch-bf16 {name:14s} max-rel {float((t - r).abs().max() / r.abs().max()):.5f}")
print("torch-bf16 final max-rel", float((wb - want).abs().max() / want.abs().max()),
      " rms-rel ours", float((got - want).norm() / want.norm()), " rms-rel torch-bf16", float((wb - want).norm() / want.norm()))
