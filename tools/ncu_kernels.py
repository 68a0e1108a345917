"""Summarise an ncu CSV launch list (any of gpu__time_duration.sum / dram__bytes_read.sum / dram__bytes_write.sum):
per-kernel launches, total device time, share, and DRAM traffic, over the LAST `n` launches (default: all).

    python tools/ncu_kernels.py gpurun_out/launches.csv [last_n]
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
last_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
launches = collections.OrderedDict()       # ID -> {name, metrics}
for r in rows:
    d = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r.get("Grid Size", "")})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = r.get("Metric Unit", "")
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)   # -> us
    elif name.startswith("dram__bytes"):
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    d[name] = v
items = list(launches.values())
if last_n:
    items = items[-last_n:]


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    n = n.replace("(anonymous namespace)::", "")
    return n[:64]


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in items:
    a = agg[short(d["name"])]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values()) or 1.0
print(f"{'kernel':64s} {'n':>5s} {'total us':>10s} {'share':>6s} {'avg us':>8s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:64s} {a[0]:5d} {a[1]:10.1f} {100 * a[1] / tot:5.1f}% {a[1] / a[0]:8.1f} {a[2] / 1e6:11.1f} {a[3] / 1e6:11.1f}")
print(f"total {tot:.1f} us over {len(items)} launches (cold-cache, serialised under ncu: compare shares, not absolutes)")
