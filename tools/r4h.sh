# third session of round 2: 8-bit AdamW v4 (one compare after the table, sign of a code from its index): parity, then time
timeout 60 python -m pytest tests/test_optim8bit_gpu.py -q -m gpu 2>&1 | tail -6 > gpurun_out/r02c_adam8bit_v4.txt
timeout 30 python tools/adam8bit_time.py >> gpurun_out/r02c_adam8bit_v4.txt 2>&1
cat gpurun_out/r02c_adam8bit_v4.txt
