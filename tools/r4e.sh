# third session of round 2: optimizer kernels at the saturating size (fp32-moment AdamW vs block-wise 8-bit AdamW) and the
# config-3 fine-tuning step with the 8-bit optimizer state (the reference's default, finetune_sd.py:300)
timeout 100 python -c "
import json, torch, bench
r = bench.elementwise_gbs(torch.device('cuda:0'))
print(json.dumps({'peak_gbs': r['peak_gbs'], 'saturating_64Mi': r['sizes']['saturating_64Mi']}))
" > gpurun_out/r02c_elementwise_with_optimizers.json 2> gpurun_out/r4e.err
timeout 200 python bench.py --workload train --optim-bits 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_train_n1_adam8bit.json 2>> gpurun_out/r4e.err
tail -c 1500 gpurun_out/r02c_elementwise_with_optimizers.json; echo; head -c 700 gpurun_out/r02c_bench_train_n1_adam8bit.json; echo; tail -3 gpurun_out/r4e.err
