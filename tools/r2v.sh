for v in "" "B200SD_PAIR=0"; do
env $v timeout 900 python bench.py --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-train-legs --no-elementwise --dump-ops "gpurun_out/r2v_ops_b8_$v.txt" > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2v_bench.json')); print('[$v]', round(d['value']/50,3), 'images/s', round(d['ms_per_step'],3), {k:(v['ms_per_step'], v.get('tflops')) if isinstance(v,dict) else v for k,v in d['kernels'].items()})"
done
head -40 "gpurun_out/r2v_ops_b8_.txt"
