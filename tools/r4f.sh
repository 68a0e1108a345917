# third session of round 2, final validation at HEAD: optimizer kernels at the saturating size (8-bit AdamW v2: lookup-table
# nearest-code search, reciprocal scaling, occupancy-sized grid), then the whole GPU suite
timeout 60 python -c "
import json, torch, bench
r = bench.elementwise_gbs(torch.device('cuda:0'))
print(json.dumps({'peak_gbs': r['peak_gbs'], 'saturating_64Mi': r['sizes']['saturating_64Mi']}))
" > gpurun_out/r02c_elementwise_with_optimizers.json 2> gpurun_out/r4f.err
python - <<'PY'
import json
try:
    k = json.load(open('gpurun_out/r02c_elementwise_with_optimizers.json'))['saturating_64Mi']['kernels']
    for n in ('adamw_step', 'adamw8bit_step'):
        print(n, k[n]['us'], 'us', k[n]['gbs'], 'GB/s', k[n]['frac_of_hbm_peak'])
except Exception as e:
    print('elementwise record FAILED', e)
PY
timeout 250 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02c_gpu_tests.txt; cat gpurun_out/r02c_gpu_tests.txt
