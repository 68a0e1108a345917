timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json')); print('persist v3', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"
B200SD_PERSIST=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise > gpurun_out/r2g_bench_old.json 2> gpurun_out/r2g_bench_old.err; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_old.json')); print('old', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])"
