# round 2, call B: persistent GEMM kernel -- parity first (bounded), then A/B bench
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_blocks_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/r2b_tests_gemm.txt
cat gpurun_out/r2b_tests_gemm.txt | tail -12
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_train_gpu.py -q 2>&1 | tail -30 > gpurun_out/r2b_tests_unet.txt
cat gpurun_out/r2b_tests_unet.txt | tail -12
timeout 300 python tools/debug_dropin.py > gpurun_out/r2b_dropin.txt 2>&1; tail -40 gpurun_out/r2b_dropin.txt
B200SD_PERSIST=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise --dump-ops gpurun_out/r2b_ops_old.txt > gpurun_out/r2b_bench_old.json 2> gpurun_out/r2b_bench_old.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise --dump-ops gpurun_out/r2b_ops_persist.txt > gpurun_out/r2b_bench_persist.json 2> gpurun_out/r2b_bench_persist.err
for f in old persist; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2b_bench_$f.json')); print('$f', round(d['value'],1), round(d['ms_per_step'],3), {k:v['ms_per_step'] for k,v in d['kernels'].items() if isinstance(v,dict)})
except Exception as e: print('$f', 'FAILED', e)
PY
done
tail -3 gpurun_out/r2b_bench_persist.err
