for v in 0 1; do
  B200SD_PIPELINED_ADAMW=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$v bench.py --workload train --gpus 8 --steps 10 --warmup 3 2>/dev/null > gpurun_out/r02b_train_n8_pipelined$v.json
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02b_train_n8_pipelined$v.json').read().strip().splitlines()[-1]); print('PIPELINED=$v N=8 train ms/step', round(d['ms_per_step'],2), 'samples/s', round(d['value'],1))"
done
