"""One eager (no CUDA graph) VAE decode of a 4x64x64 latent on random-init SD v1.x weights: the target of ncu captures."""
import sys, torch
sys.path.insert(0, '.')
from b200sd.vae import AutoencoderKL
torch.manual_seed(0)
vae = AutoencoderKL().to('cuda:0').eval()
vae.use_cuda_graph = False
z = torch.randn(1, 4, 64, 64, device='cuda:0')
for _ in range(2):
    img = vae.decode(z).sample
torch.cuda.synchronize()
print(tuple(img.shape), float(img.abs().mean()))
