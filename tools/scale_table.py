"""Markdown tables of the round's multi-GPU records: profiles/r02_bench_n{1,2,4,8}.json (default bench line, incl. the fine-tuning
legs) and profiles/r02_sweep_512x768_n{1,2,4,8}.jsonl (config 5)."""
import json, os, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")

def last_json(path):
    return [json.loads(l) for l in open(path) if l.lstrip().startswith("{")]

rows = {}
for n in (1, 2, 4, 8):
    f = os.path.join(P, f"{R}_bench_n{n}.json")
    if os.path.exists(f):
        rows[n] = last_json(f)[-1]
if rows:
    b = rows[min(rows)]
    print("| GPUs | sampling it/s (x) | e2e it/s | UNet fine-tune ms/step, samples/s (x), exposed comm | TE fine-tune ms/step, samples/s (x), exposed comm |")
    print("|---|---|---|---|---|")
    for n, d in rows.items():
        t, tt = d["train"], d["train_text"]
        print(f"| {n} | {d['value']:.1f} ({d['value'] / b['value']:.2f}x) | {d['e2e']['value']:.1f} | {t['ms_per_step']:.1f}, {t['samples_per_s']:.0f} "
              f"({t['samples_per_s'] / b['train']['samples_per_s']:.2f}x), {t['exposed_comm_ms']:.1f} ms | {tt['ms_per_step']:.1f}, {tt['samples_per_s']:.0f} "
              f"({tt['samples_per_s'] / b['train_text']['samples_per_s']:.2f}x), {tt['exposed_comm_ms']:.1f} ms |")
sw = {}
for n in (1, 2, 4, 8):
    f = os.path.join(P, f"{R}_sweep_512x768_n{n}.jsonl")
    if os.path.exists(f):
        sw[n] = {d["total_images"]: d for d in last_json(f)}
if sw:
    bs = sorted({b for v in sw.values() for b in v})
    print()
    print("| images (512x768) | " + " | ".join(f"{n} GPU{'s' if n > 1 else ''}" for n in sw) + " |")
    print("|---|" + "---|" * len(sw))
    for b in bs:
        cells = []
        for n, v in sw.items():
            d = v.get(b)
            cells.append("-" if d is None else f"{d['value']:.2f}" + (f" ({d['gpus_idle']} idle)" if d["gpus_idle"] else ""))
        print(f"| {b} | " + " | ".join(cells) + " |")
