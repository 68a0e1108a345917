python -m pytest tests/test_unet_gpu.py -q -m gpu -k "plms or PLMS or captured" 2>&1 | tail -6
