python tools/one_attn.py 2 4096 40 | tail -1
B200SD_ATTN_SMEM_PAD=80000 python tools/one_attn.py 2 4096 40 | tail -1 | sed 's/^/1 CTA per SM: /'
python tools/one_attn.py 8 4096 40 | tail -1
B200SD_ATTN_SMEM_PAD=80000 python tools/one_attn.py 8 4096 40 | tail -1 | sed 's/^/1 CTA per SM: /'
