# ncu full sets of the GroupNorm / LayerNorm kernels and of a split-K small-M conv, inside the real (eager) forward
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on --kernel-name regex:gn_parts_kernel --launch-skip 70 --launch-count 2 -o gpurun_out/r2o_gn python tools/profile_step.py > gpurun_out/r2o_gn.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on --kernel-name regex:layernorm_kernel --launch-skip 50 --launch-count 1 -o gpurun_out/r2o_ln python tools/profile_step.py > gpurun_out/r2o_ln.log 2>&1
ls -la gpurun_out/r2o_*.ncu-rep
