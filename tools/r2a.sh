# round 2, call A: GPU suite at HEAD, then the sampling bench with row-major vs k-block-major weights and with PDL
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2a_tests.txt
B200SD_W_KMAJOR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise --dump-ops gpurun_out/r2a_ops_rowmajor.txt > gpurun_out/r2a_bench_rowmajor.json 2> gpurun_out/r2a_bench_rowmajor.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-ops gpurun_out/r2a_ops_kmajor.txt > gpurun_out/r2a_bench_kmajor.json 2> gpurun_out/r2a_bench_kmajor.err
B200SD_PDL=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise --dump-ops gpurun_out/r2a_ops_pdl.txt > gpurun_out/r2a_bench_pdl.json 2> gpurun_out/r2a_bench_pdl.err
cat gpurun_out/r2a_tests.txt
for f in rowmajor kmajor pdl; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2a_bench_$f.json')); print('$f', round(d['value'],1), round(d['ms_per_step'],3), {k:v['ms_per_step'] for k,v in d['kernels'].items() if isinstance(v,dict)})
except Exception as e: print('$f', 'FAILED', e)
PY
done
tail -3 gpurun_out/r2a_bench_kmajor.err
