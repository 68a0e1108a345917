python -m pytest tests/test_dropin_gpu.py tests/test_clip_gpu.py -q -m gpu -k "reference or four" 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --workload sweep --portrait --sweep-batches 1,4,16 --steps 10 --warmup 3 2>/dev/null | cut -c1-150
