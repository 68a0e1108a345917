timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_blocks_gpu.py -q 2>&1 | tail -2
for v in "" "B200SD_PERSIST=0"; do
env $v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise --dump-ops gpurun_out/r2k_ops_$v.txt > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print('[$v]', round(d['value'],2), round(d['ms_per_step'],4), {k:(v['ms_per_step'] if isinstance(v,dict) else v) for k,v in d['kernels'].items()})"
done
python - <<'PY'
import re,collections
def load(f):
    d=collections.defaultdict(list)
    for l in open(f):
        if l.startswith('#'): continue
        m=re.match(r'\s*([\d.]+) us\s+([\d.]+) TF/s\s+(.*)',l)
        d[m.group(3).strip()].append(float(m.group(1)))
    return d
a=load('gpurun_out/r2k_ops_B200SD_PERSIST=0.txt'); b=load('gpurun_out/r2k_ops_.txt')
rows=[]
for k in a:
    sa=sum(a[k]); sb=sum(b.get(k,[0]))
    rows.append((sb-sa,k,len(a[k]),sa/len(a[k]),sb/max(1,len(b.get(k,[1])))))
rows.sort()
for d,k,n,x,y in rows:
    if abs(d)>4: print(f"{k:40s} n={n:3d} old {x:7.1f} us  persist {y:7.1f} us  total delta {d:8.1f} us")
PY
