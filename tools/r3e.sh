python -m pytest tests/test_vae_gpu.py tests/test_blocks_gpu.py -q -m gpu 2>&1 | tail -4
python tools/vae_profile.py 1 > gpurun_out/r3e_vae_profile.txt 2>&1; head -24 gpurun_out/r3e_vae_profile.txt
python tools/vae_profile.py 4 > gpurun_out/r3e_vae_profile_b4.txt 2>&1; head -8 gpurun_out/r3e_vae_profile_b4.txt
