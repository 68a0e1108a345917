python -m pytest tests/test_clip_gpu.py -x -q -m gpu > gpurun_out/r3b_tests.txt 2>&1; tail -15 gpurun_out/r3b_tests.txt
python bench.py --workload train_text --steps 8 --warmup 3 > gpurun_out/r3b_train_text.json 2> gpurun_out/r3b_train_text.err || tail -20 gpurun_out/r3b_train_text.err
cut -c1-400 gpurun_out/r3b_train_text.json
