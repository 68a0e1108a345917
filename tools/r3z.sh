# session-2 (r02b) round-end validation at HEAD
R=r02b bash tools/final_validation.sh 2>&1 | tail -30
# ncu: per-launch time + DRAM bytes of one eager VAE decode, and one full-set capture of its widest conv (rows wider than a tile)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02b_ncu_vae_decode.csv python tools/vae_decode_once.py > gpurun_out/r02b_ncu_vae.log 2>&1
python tools/ncu_kernels.py gpurun_out/r02b_ncu_vae_decode.csv 77 > gpurun_out/r02b_ncu_vae_decode_summary.txt 2>&1; head -20 gpurun_out/r02b_ncu_vae_decode_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_persist_kernel --launch-skip 50 --launch-count 2 -o gpurun_out/r02b_vae_conv python tools/vae_decode_once.py > gpurun_out/r02b_ncu_vae_full.log 2>&1
ncu -i gpurun_out/r02b_vae_conv.ncu-rep --page raw --csv > gpurun_out/r02b_vae_conv_raw.csv 2>/dev/null
ls -la gpurun_out/r02b_vae_conv.ncu-rep
