timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_tcgen05_kernel --launch-skip 12 --launch-count 3 -o gpurun_out/r02b_vae_conv python tools/vae_decode_once.py > gpurun_out/r02b_ncu_vae_full.log 2>&1
ncu -i gpurun_out/r02b_vae_conv.ncu-rep --page raw --csv > gpurun_out/r02b_vae_conv_raw.csv 2>/dev/null
ls -la gpurun_out/r02b_vae_conv.ncu-rep; tail -3 gpurun_out/r02b_ncu_vae_full.log
