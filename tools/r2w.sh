timeout 900 python -m pytest tests/test_unet_gpu.py -q -s -k "large_batch or config1" 2>&1 | grep -E "passed|failed|FAILED|max-rel|Error" | tail -6
for v in "" "B200SD_GN_CTAS_PER_SM=4" "B200SD_GN_CTAS_PER_SM=6"; do
echo "== [$v]"
env $v timeout 900 python bench.py --workload sweep --sweep-batches 1,8,16 --steps 10 --warmup 3 2> gpurun_out/r2w.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['total_images'], round(d['value'],3), 'images/s', round(d['ms_per_iteration_max_over_ranks'],3), 'ms/it', round(d['tflops_per_active_gpu'],1), 'TF/s')"
done
tail -3 gpurun_out/r2w.err
