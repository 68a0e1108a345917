N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02b_bench_n$N.json 2> gpurun_out/r02b_bench_n$N.err; tail -2 gpurun_out/r02b_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02b_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
print("train", d["train"]["ms_per_step"], d["train"]["samples_per_s"], d["train"]["exposed_comm_ms"], "train_text", d["train_text"]["ms_per_step"], d["train_text"]["samples_per_s"], d["train_text"]["exposed_comm_ms"])
PY
