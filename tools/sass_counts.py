"""Per-kernel SASS mnemonic counts of libb200sd.so (what proves a Blackwell-native kernel, B200_PROFILING.md):
    python tools/sass_counts.py > profiles/r02_sass_mnemonics.txt
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, HMMA = legacy mma.sync."""
import collections, os, re, subprocess, sys
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stable-diffusion-for-book-cover-generation_b200", "libb200sd.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "HMMA", "MUFU.EX2", "LDGSTS", "SYNCS", "REDG", "RED."]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in keys:
        if re.search(r"(?<![A-Z])" + re.escape(k), line):     # "HMMA" must not count "UTCHMMA"
            counts[cur][k] += 1
    counts[cur]["_instr"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in keys) + "   instr")
for (name, c), dn in zip(counts.items(), demangle):
    dn = re.sub(r"\(anonymous namespace\)::", "", dn)
    dn = re.sub(r"\(.*$", "", dn).replace("void ", "")
    print(f"{dn[:70]:70s} " + " ".join(f"{c[k]:8d}" for k in keys) + f" {c['_instr']:7d}")
