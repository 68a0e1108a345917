for v in "" "B200SD_PERSIST=0" "B200SD_PERSIST_MASK=1"; do
echo "== [$v]"
env $v timeout 900 python bench.py --workload sweep --sweep-batches 4,8,16 --steps 10 --warmup 3 2> gpurun_out/r2u.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['total_images'], round(d['value'],3), 'images/s', round(d['ms_per_iteration_max_over_ranks'],2), 'ms/it', round(d['tflops_per_active_gpu'],1), 'TF/s')"
done
tail -3 gpurun_out/r2u.err
