# round 2, session 2: captured sampler + lanes
python -m pytest tests/test_unet_gpu.py -x -q -m gpu -k "captured or denoise_loop_takes" > gpurun_out/r3a_tests.txt 2>&1; tail -5 gpurun_out/r3a_tests.txt
F="--no-train-legs --no-cpu-baseline --no-elementwise --steps 50 --warmup 5"
for cfg in "1 1" "1 2" "2 1" "2 2" "2 4" "4 1" "4 2" "8 1" "8 2"; do
  set -- $cfg
  python bench.py $F --batch $1 --lanes $2 > gpurun_out/r3a_b$1_l$2.json 2> gpurun_out/r3a_b$1_l$2.err || tail -5 gpurun_out/r3a_b$1_l$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3a_b$1_l$2.json").read().strip().splitlines()[-1])
    print("B=$1 lanes=$2", round(d["ms_per_step"],3), "ms/step", round(d["images_per_s"],3), "img/s  e2e", round(d["e2e"]["value"],1), "it/s  roof", round(d["roofline"]["frac"],3) if d["roofline"] else None)
except Exception as e: print("B=$1 lanes=$2 failed", e)
PY
done
