python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02b_gpu_tests.txt; cat gpurun_out/r02b_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.txt 2>&1; echo "smoke rc $?" >> gpurun_out/r02b_smoke.txt; tail -2 gpurun_out/r02b_smoke.txt
python bench.py --dump-ops gpurun_out/r02b_in_step_per_launch_times.txt > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; wc -l gpurun_out/r02b_bench_n1.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02b_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], "t2i", d["text_to_image"]["1_images"]["images_per_s"], d["text_to_image"]["4_images"]["images_per_s"], "train", d["train"]["ms_per_step"], d["train_text"]["ms_per_step"])
PY
