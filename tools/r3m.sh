timeout 900 python -m pytest tests/test_train_gpu.py tests/test_dropin_gpu.py tests/test_clip_gpu.py -q -m gpu -x 2>&1 | tail -4
for v in 0 1; do
  B200SD_TRAIN_GN_FROM_GEMM=$v python bench.py --workload train --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('GN_FROM_GEMM=$v train ms/step', round(d['ms_per_step'],2))"
done
python bench.py --workload train_text --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train_text ms/step', round(d['ms_per_step'],2))"
