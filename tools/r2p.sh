timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_blocks_gpu.py -q -x 2>&1 | tail -12
SMALL_ONLY=1 timeout 600 python tools/cold_probe.py 2>&1 | grep -v Warn | cut -c1-200
timeout 600 python -m pytest tests/test_unet_gpu.py -q 2>&1 | tail -3
