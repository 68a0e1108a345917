"""Join an ncu `gpu__time_duration.sum` launch list with the plan-entry names written by tools/profile_step.py (the LAST
len(names) launches of the CSV are the named forward).  Prints per-kind totals and the per-shape table.

    python tools/ncu_join.py launches.csv names.txt [out.txt]
"""
import collections, csv, sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
rows = [r for r in rows if r["Metric Name"] == "gpu__time_duration.sum"]
names = [l.rstrip("\n").split("\t") for l in open(sys.argv[2]) if l.strip()]
last = rows[-len(names):]
assert len(last) == len(names), (len(last), len(names))
out = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout


def us(r):
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "ns")
    return v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)


kinds = collections.defaultdict(lambda: [0, 0.0, 0.0])
shapes = collections.defaultdict(lambda: [0, 0.0, 0.0, ""])
for r, (kind, name, flops) in zip(last, names):
    t = us(r)
    k = kinds[kind]; k[0] += 1; k[1] += t; k[2] += float(flops)
    s = shapes[name]; s[0] += 1; s[1] += t; s[2] += float(flops); s[3] = r["Kernel Name"].split("(")[0][-40:]
tot = sum(k[1] for k in kinds.values())
print(f"# {len(names)} launches, sum of kernel durations {tot:.1f} us (ncu gpu__time_duration.sum: serialised, no launch gaps)", file=out)
for kind, (n, t, fl) in sorted(kinds.items(), key=lambda kv: -kv[1][1]):
    print(f"{kind:12s} n={n:4d} {t:9.1f} us {100 * t / tot:5.1f}%  {fl / t / 1e6 if fl else 0:7.1f} TF/s", file=out)
print(file=out)
for name, (n, t, fl, kn) in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us total  n={n:3d} avg {t / n:7.2f} us  {fl / t / 1e6 if fl else 0:7.1f} TF/s  {name:36s} {kn}", file=out)
