for v in ${POLYS:-0 4 3 2}; do
  for shp in "2 4096 40" "8 4096 40" "8 1024 80" "2 4096 40 6"; do
    B200SD_ATTN_POLY=$v timeout 120 python tools/one_attn.py $shp 2>&1 | tail -1 | sed "s/^/poly $v /"
  done
done
