python -m pytest tests/test_train_gpu.py -q -m gpu -k "adamw or trainer" 2>&1 | tail -3
for v in 0 1; do
  B200SD_PIPELINED_ADAMW=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2960$v bench.py --workload train --gpus 2 --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PIPELINED=$v N=2 train ms/step', round(d['ms_per_step'],2), 'loss', d['config'].get('last_loss'))"
done
