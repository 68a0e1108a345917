timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_blocks_gpu.py -q 2>&1 | tail -2
timeout 300 python tools/conv_probe.py 2>&1 | cut -c1-215
for v in "" "B200SD_PERSIST=0"; do
env $v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench.json')); print('[$v]', round(d['value'],2), round(d['ms_per_step'],4), {k:(v['ms_per_step'] if isinstance(v,dict) else v) for k,v in d['kernels'].items()})"
done
