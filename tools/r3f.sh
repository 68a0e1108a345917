python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r3f_bench_n2.json 2> gpurun_out/r3f_bench_n2.err; tail -3 gpurun_out/r3f_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3f_bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
print("train", d["train"]["ms_per_step"], d["train"]["exposed_comm_ms"], "train_text", d["train_text"]["ms_per_step"], d["train_text"]["exposed_comm_ms"])
PY
