python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02b_gpu_tests.txt; cat gpurun_out/r02b_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.txt 2>&1; echo "smoke rc $?" >> gpurun_out/r02b_smoke.txt; tail -2 gpurun_out/r02b_smoke.txt
