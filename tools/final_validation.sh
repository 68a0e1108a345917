set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1; echo "smoke rc $?" >> gpurun_out/final_smoke.txt
python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/bench.err
python bench.py --workload train > gpurun_out/r01_bench_train_n1.json 2>> gpurun_out/bench.err
python bench.py --precision fp32 > gpurun_out/r01_bench_n1_fp32_path.json 2>> gpurun_out/bench.err
python tools/per_launch.py > gpurun_out/r01_warm_per_launch_times.txt 2>> gpurun_out/bench.err
ITERS=2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01_ncu_unet_forward.csv python tools/profile_step.py > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/final_tests.txt; tail -3 gpurun_out/final_smoke.txt; tail -2 gpurun_out/ncu_step.log
