timeout 300 python tools/gemm_trace.py 2>&1 | grep -v Warn > gpurun_out/r2f_trace.txt; cat gpurun_out/r2f_trace.txt
timeout 300 python tools/conv_probe.py 2>&1 | grep "geglu\|M8192 N320 K320\|N960" 
timeout 300 python -m pytest tests/test_gemm_gpu.py -q 2>&1 | tail -3
