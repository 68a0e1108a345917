timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:attention_tc_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02b_attn_fwd python tools/one_attn.py 2 4096 40 > gpurun_out/r02b_ncu_attn.log 2>&1
ncu -i gpurun_out/r02b_attn_fwd.ncu-rep --page raw --csv > gpurun_out/r02b_attn_fwd_raw.csv 2>/dev/null
ncu -i gpurun_out/r02b_attn_fwd.ncu-rep --page source --csv > gpurun_out/r02b_attn_fwd_source.csv 2>/dev/null
ls -la gpurun_out/r02b_attn_fwd* | cut -c1-100
