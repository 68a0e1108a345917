"""Summarise an ncu `gpu__time_duration.sum` launch list (CSV) of tools/profile_step.py:
per-kernel totals for the LAST forward, and the GEMM kernel broken down by grid shape."""
import collections, csv, re, sys

path = sys.argv[1]
per_fwd = int(sys.argv[2]) if len(sys.argv) > 2 else 406
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
pat = re.compile(r'(gemm_tcgen05_kernel|attention_kernel<\d+>|gn_stats_kernel|gn_apply_kernel|layernorm_kernel<\d+>|'
                 r'small_linear_kernel|temb_kernel|conv_in_kernel|conv_out_kernel|upsample2x_kernel|im2col_s2_kernel|'
                 r'cfg_ddim_kernel|cfg_plms_kernel|add_noise_kernel|mse_\w+_kernel|\w+_kernel)')
def short(n):
    m = pat.search(n)
    return m.group(1) if m else None
ours = [x for x in rows if short(x['Kernel Name'])]
last = ours[-per_fwd:]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for x in last:
    n = short(x['Kernel Name']); v = float(x['Metric Value'].replace(',', '')) / 1e3
    agg[n][0] += 1; agg[n][1] += v; agg[n][2] = max(agg[n][2], v)
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':28s} {'n':>4s} {'total us':>10s} {'share':>6s} {'avg us':>8s} {'max us':>8s}")
for n, (c, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:28s} {c:4d} {t:10.1f} {100 * t / tot:5.1f}% {t / c:8.1f} {mx:8.1f}")
print(f"total {tot:.1f} us over {len(last)} launches")
byg = collections.defaultdict(lambda: [0, 0.0])
for x in last:
    if 'gemm_tcgen05' in x['Kernel Name']:
        v = float(x['Metric Value'].replace(',', '')) / 1e3
        byg[x['Grid Size']][0] += 1; byg[x['Grid Size']][1] += v
print("\ngemm_tcgen05_kernel by grid (m_tiles, n_tiles, split_k):")
for gs, (c, t) in sorted(byg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {gs:16s} n={c:3d} avg {t / c:7.1f} us  total {t:8.1f} us")
