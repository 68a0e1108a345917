python bench.py --dump-ops gpurun_out/r02b_in_step_per_launch_times.txt > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; wc -l gpurun_out/r02b_bench_n1.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02b_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], "t2i", d["text_to_image"]["1_images"]["images_per_s"], d["text_to_image"]["4_images"]["images_per_s"], "train", d["train"]["ms_per_step"], d["train_text"]["ms_per_step"])
PY
python -m pytest tests/test_train_gpu.py -q -m gpu 2>&1 | tail -2
