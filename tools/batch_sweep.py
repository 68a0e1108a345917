"""BASELINE config 5 style sweep: images/s of 50-step CFG DDIM sampling vs images per GPU (UNet batch = 2x).
   python tools/batch_sweep.py [--portrait] 1 2 4 8 16"""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
portrait = "--portrait" in sys.argv
batches = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2, 4, 8, 16]
for b in batches:
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--steps", "20", "--warmup", "3", "--batch", str(b), "--no-cpu-baseline"] + (["--portrait"] if portrait else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        print(f"batch {b}: failed: {r.stderr[-300:]}")
        continue
    k = d["kernels"]
    print(json.dumps({"images_per_gpu": b, "latent": d["config"]["latent"], "it_per_s": round(d["value"], 2), "ms_per_step": round(d["ms_per_step"], 3),
                      "images_per_s": round(d["config"]["images_per_s"], 3), "e2e_images_per_s": round(d["config"]["e2e_images_per_s"], 3),
                      "tflops_end_to_end": round(d["config"]["tflops_end_to_end"], 1), "gemm_conv_tflops": round(d["roofline"]["achieved"], 1),
                      "gemm_frac_of_peak": round(d["roofline"]["frac"], 3),
                      "kernel_ms": {n: v["ms_per_step"] for n, v in k.items()}}), flush=True)
