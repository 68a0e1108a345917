for P in 4 3 2 0; do
  for cfg in "2 4096 40" "2 1024 80" "8 4096 40"; do
    B200SD_ATTN_POLY=$P python tools/one_attn.py $cfg 2>&1 | tail -1 | sed "s/^/poly=$P /"
  done
done
