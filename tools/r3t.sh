for W in 4 8; do
  for cfg in "2 4096 40" "8 4096 40" "2 1024 80" "2 256 160" "2 4096 40 3.0"; do
    B200SD_ATTN_WARPS=$W timeout 120 python tools/one_attn.py $cfg 2>&1 | tail -1 | sed "s/^/warps=$W /"
  done
done
timeout 600 python -m pytest tests/test_blocks_gpu.py -q -m gpu -k "attention" 2>&1 | tail -3
