timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_blocks_gpu.py -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_unet_gpu.py -q -s 2>&1 | grep -E "passed|failed|FAILED|max-rel|cosine|seeds" | tail -12
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2s_bench.json')); print(round(d['value'],2), round(d['ms_per_step'],4), {k:(v['ms_per_step'] if isinstance(v,dict) else v) for k,v in d['kernels'].items() if k in ('gemm','conv3x3','groupnorm', 'attention','layernorm')})"
timeout 900 python -m pytest tests/test_train_gpu.py -q -s -k "batch8" 2>&1 | grep -E "passed|failed|FAILED|max-rel|cosine" | tail -6
