timeout 600 python -m pytest tests/test_unet_gpu.py -q 2>&1 | tail -3
for v in "" "B200SD_PREFETCH=0" "B200SD_PSPLIT=0"; do
env $v timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2q_bench.json')); print('[$v]', round(d['value'],2), round(d['ms_per_step'],4), {k:(v['ms_per_step'] if isinstance(v,dict) else v) for k,v in d['kernels'].items() if k in ('gemm','conv3x3','groupnorm', 'attention')})"
done
