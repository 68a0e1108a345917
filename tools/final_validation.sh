# Round-end validation at HEAD: the GPU suite, smoke, the default bench line (sampling + fine-tuning legs + elementwise GB/s + cpu baseline),
# the reference arm, per-shape ncu launch list of one forward, and the SASS mnemonic table.
R=${R:-r02}
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${R}_gpu_tests.txt; cat gpurun_out/${R}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.txt 2>&1; echo "smoke rc $?" >> gpurun_out/${R}_smoke.txt; tail -2 gpurun_out/${R}_smoke.txt
python bench.py --dump-ops gpurun_out/${R}_in_step_per_launch_times.txt > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference_cpu.json 2>> gpurun_out/${R}_bench_n1.err
python bench.py --impl library > gpurun_out/${R}_bench_library_bf16.json 2>> gpurun_out/${R}_bench_n1.err
python bench.py --precision fp32 --no-cpu-baseline > gpurun_out/${R}_bench_n1_fp32_path.json 2>> gpurun_out/${R}_bench_n1.err
python bench.py --workload sweep --sweep-batches 1,2,4,8,16,32 --steps 10 --warmup 3 > gpurun_out/${R}_sweep_512px_n1.jsonl 2>> gpurun_out/${R}_bench_n1.err
python bench.py --workload sweep --portrait --steps 10 --warmup 3 > gpurun_out/${R}_sweep_512x768_n1.jsonl 2>> gpurun_out/${R}_bench_n1.err
python - <<PY
import json
for f in ('${R}_bench_n1','${R}_bench_reference_cpu','${R}_bench_library_bf16','${R}_bench_n1_fp32_path'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['value'],3), d.get('unit'), round(d['ms_per_step'],3), (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(f, 'FAILED', e)
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none --csv --log-file gpurun_out/${R}_ncu_unet_forward.csv python tools/profile_step.py gpurun_out/${R}_ncu_names.txt > gpurun_out/${R}_ncu.log 2>&1
python tools/ncu_join.py gpurun_out/${R}_ncu_unet_forward.csv gpurun_out/${R}_ncu_names.txt gpurun_out/${R}_ncu_unet_forward_summary.txt; head -12 gpurun_out/${R}_ncu_unet_forward_summary.txt
tail -3 gpurun_out/${R}_bench_n1.err
