python tools/vae_profile.py 1 > gpurun_out/r3d_vae_profile.txt 2>&1; cat gpurun_out/r3d_vae_profile.txt | head -60
python -m pytest tests/test_vae_gpu.py -q -m gpu -k "pipeline" 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/r3d_bench_n1.json 2> gpurun_out/r3d_bench_n1.err || tail -20 gpurun_out/r3d_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3d_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"])
print("t2i", d["text_to_image"])
print("train", d["train"]["ms_per_step"], "train_text", d["train_text"]["ms_per_step"])
PY
