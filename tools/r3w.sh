python tools/one_wide_conv.py 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_ --launch-skip 12 --launch-count 1 -o gpurun_out/r02b_vae_conv_wide python tools/one_wide_conv.py > gpurun_out/r02b_ncu_vae_full2.log 2>&1
ncu -i gpurun_out/r02b_vae_conv_wide.ncu-rep --page raw --csv > gpurun_out/r02b_vae_conv_wide_raw.csv 2>/dev/null
ls -la gpurun_out/r02b_vae_conv_wide.ncu-rep | cut -c1-80
