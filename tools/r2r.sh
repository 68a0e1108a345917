for v in "B200SD_W_KMAJOR=1 B200SD_PREFETCH=0" "B200SD_W_KMAJOR=1 B200SD_PREFETCH=4" "B200SD_W_KMAJOR=1 B200SD_PREFETCH=8" "B200SD_W_KMAJOR=1 B200SD_PREFETCH=16" "B200SD_PREFETCH=0"; do
env $v timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2r_bench.json')); print('[$v]', round(d['value'],2), round(d['ms_per_step'],4), {k:(v['ms_per_step'] if isinstance(v,dict) else v) for k,v in d['kernels'].items() if k in ('gemm','conv3x3','groupnorm', 'attention')})"
done
