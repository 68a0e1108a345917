for v in 0 1 2 3; do
  for shp in "2 4096 40" "2 1024 80" "8 4096 40" "2 4096 40 6" "2 1024 80 6"; do
    B200SD_ATTN_FWD=$v timeout 120 python tools/one_attn.py $shp 2>&1 | tail -1
  done
done
