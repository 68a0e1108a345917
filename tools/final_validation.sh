# Round-end validation: the GPU suite (unet tests right after the block tests -- the order that once exposed an intermittent
# failure -- then everything else), smoke, and the bench lines that go to profiles/.
python -m pytest tests/test_blocks_gpu.py tests/test_elementwise_gpu.py tests/test_unet_gpu.py -q 2>&1 | grep -E "cosine|passed|failed" > gpurun_out/final_tests.txt
python -m pytest tests/test_gemm_gpu.py tests/test_gemm_bwd_gpu.py tests/test_backward_ops_gpu.py tests/test_train_gpu.py -q 2>&1 | grep -E "passed|failed|FAILED" >> gpurun_out/final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1; echo "smoke rc $?" >> gpurun_out/final_smoke.txt
python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/bench.err
python bench.py --workload train > gpurun_out/r01_bench_train_n1.json 2>> gpurun_out/bench.err
cat gpurun_out/final_tests.txt; tail -2 gpurun_out/final_smoke.txt
python -c "
import json
for f in ('r01_bench_n1','r01_bench_train_n1'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']['reasons'])
"
