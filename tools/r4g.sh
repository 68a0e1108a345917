# third session of round 2: 8-bit AdamW v3 (4 values per thread, 512-thread CTAs), compiled for 3 and for 4 resident CTAs per SM:
# bit-exact parity of both against the oracle, then the time of each
timeout 60 python -m pytest tests/test_optim8bit_gpu.py -q -m gpu 2>&1 | tail -4 > gpurun_out/r02c_adam8bit_variants.txt
B200SD_ADAM8_CTAS=4 timeout 40 python -m pytest tests/test_optim8bit_gpu.py -q -m gpu -k bit_exact 2>&1 | tail -4 >> gpurun_out/r02c_adam8bit_variants.txt
B200SD_ADAM8_CTAS=3 timeout 30 python tools/adam8bit_time.py >> gpurun_out/r02c_adam8bit_variants.txt 2>&1
B200SD_ADAM8_CTAS=4 timeout 30 python tools/adam8bit_time.py >> gpurun_out/r02c_adam8bit_variants.txt 2>&1
cat gpurun_out/r02c_adam8bit_variants.txt
