for m in 0 1 2 4 8 15; do
B200SD_PERSIST_MASK=$m timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_bench.json')); print('[mask $m]', round(d['value'],2), round(d['ms_per_step'],4), {k:(v['ms_per_step'] if isinstance(v,dict) else v) for k,v in d['kernels'].items() if k in ('gemm','conv3x3','groupnorm')})"
done
