# round 2, call E: probe of tile width / ring depth, GEGLU timeline, full GPU suite at HEAD
timeout 300 python tools/conv_probe.py > gpurun_out/r2e_probe.txt 2>&1; cat gpurun_out/r2e_probe.txt
timeout 300 python tools/gemm_trace.py 2>&1 | grep -A10 GEGLU > gpurun_out/r2e_trace_geglu.txt; cat gpurun_out/r2e_trace_geglu.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2e_tests.txt; tail -8 gpurun_out/r2e_tests.txt
