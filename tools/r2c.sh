# per-kernel durations under ncu: old kernel vs persistent kernel
for mode in 0 1; do
  B200SD_PERSIST=$mode timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_$mode.csv python tools/profile_step.py gpurun_out/r2c_names_$mode.txt > gpurun_out/r2c_ncu_$mode.log 2>&1
  python tools/ncu_join.py gpurun_out/r2c_launches_$mode.csv gpurun_out/r2c_names_$mode.txt gpurun_out/r2c_join_$mode.txt; head -12 gpurun_out/r2c_join_$mode.txt
done
