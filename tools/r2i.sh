timeout 300 python tools/conv_probe.py 2>&1 | grep "conv\|M8192 N320 K320\|N960" | cut -c1-200
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise --dump-ops gpurun_out/r2i_ops_persist.txt > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
B200SD_PERSIST=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train-legs --no-elementwise --dump-ops gpurun_out/r2i_ops_old.txt > gpurun_out/r2i_bench_old.json 2> gpurun_out/r2i_bench_old.err
python - <<'PY'
import re,collections
def load(f):
    d=collections.defaultdict(list)
    for l in open(f):
        if l.startswith('#'): continue
        m=re.match(r'\s*([\d.]+) us\s+([\d.]+) TF/s\s+(.*)',l)
        d[m.group(3).strip()].append(float(m.group(1)))
    return d
a=load('gpurun_out/r2i_ops_old.txt'); b=load('gpurun_out/r2i_ops_persist.txt')
rows=[]
for k in a:
    sa=sum(a[k]); sb=sum(b.get(k,[0]))
    rows.append((sb-sa,k,len(a[k]),sa/len(a[k]),sb/max(1,len(b.get(k,[1])))))
rows.sort()
for d,k,n,x,y in rows:
    if abs(d)>4: print(f"{k:40s} n={n:3d} old {x:7.1f} us  persist {y:7.1f} us  total delta {d:8.1f} us")
PY
