python bench.py --no-cpu-baseline --no-elementwise > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; echo "bench rc $?"
python -c "import torch; print('cuda available', torch.cuda.is_available()); print(torch.zeros(1).cuda())" 2>&1 | tail -3
nvidia-smi --query-gpu=name,memory.used,clocks.sm --format=csv 2>&1 | tail -2
python -m pytest tests/test_train_gpu.py -q -m gpu -rs 2>&1 | tail -8
