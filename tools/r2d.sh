# round 2, call D: persistent kernel v2 (8 epilogue warps, fast GELU): parity, timeline trace, ncu A/B (warm L2), drop-in debug
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_blocks_gpu.py -q 2>&1 | tail -15 > gpurun_out/r2d_tests_gemm.txt; tail -6 gpurun_out/r2d_tests_gemm.txt
timeout 300 python tools/gemm_trace.py > gpurun_out/r2d_trace_persist.txt 2>&1; B200SD_PERSIST=0 timeout 300 python tools/gemm_trace.py > gpurun_out/r2d_trace_old.txt 2>&1
for mode in 0 1; do
  B200SD_PERSIST=$mode timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2d_launches_$mode.csv python tools/profile_step.py gpurun_out/r2d_names_$mode.txt > gpurun_out/r2d_ncu_$mode.log 2>&1
  python tools/ncu_join.py gpurun_out/r2d_launches_$mode.csv gpurun_out/r2d_names_$mode.txt gpurun_out/r2d_join_$mode.txt; head -10 gpurun_out/r2d_join_$mode.txt
done
timeout 300 python tools/conv_probe.py > gpurun_out/r2d_probe.txt 2>&1; cat gpurun_out/r2d_probe.txt
timeout 300 python tools/debug_dropin.py > gpurun_out/r2d_dropin.txt 2>&1; grep -v Warning gpurun_out/r2d_dropin.txt | tail -30
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise > gpurun_out/r2d_bench_persist.json 2> gpurun_out/r2d_bench_persist.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_persist.json')); print('persist', d['value'], d['ms_per_step'])"
