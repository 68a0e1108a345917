python -m pytest tests/test_vae_gpu.py -q -m gpu > gpurun_out/r3c_tests.txt 2>&1; grep -E "AssertionError|passed|failed|Error:" gpurun_out/r3c_tests.txt | head -20
python - <<'PY' > gpurun_out/r3c_vae_time.txt 2>&1
import sys, torch, time
sys.path.insert(0, '.')
from b200sd.vae import AutoencoderKL
torch.manual_seed(0)
vae = AutoencoderKL().to('cuda:0').eval()
for B, h, w in ((1, 64, 64), (4, 64, 64), (1, 96, 64)):
    z = torch.randn(B, 4, h, w, device='cuda:0')
    img = torch.randn(B, 3, 8 * h, 8 * w, device='cuda:0')
    for name, fn in (("decode", lambda: vae.decode(z)), ("encode", lambda: vae.encode(img))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        eng = vae._engines[("dec" if name == "decode" else "enc", B, h if name == "decode" else 8 * h, w if name == "decode" else 8 * w, 0)]
        fl = sum(m[1] for m in eng.plan.meta)
        ms = e0.elapsed_time(e1) / 5
        print(f"{name} B={B} {8*h}x{8*w}: {ms:.2f} ms  ({fl/1e12:.2f} TFLOP, {fl/ms/1e9:.0f} TFLOP/s, {len(eng.plan)} launches, activations {eng.activation_bytes/1e9:.2f} GB)")
PY
cat gpurun_out/r3c_vae_time.txt | tail -8
