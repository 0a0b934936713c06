"""Per-kernel SASS mnemonic histogram of the shipped libmcn.so (cuobjdump -sass): the evidence that
the hot path is tcgen05 / TMEM / TMA code and contains no legacy HMMA and no floating-point atomics.
Usage: python scripts/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "myconvnet_b200", "libmcn.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "LDGSTS", "ELECT",
        "HMMA", "ATOMG", "ATOMS", "RED", "REDG", "LDG", "STG", "LDS", "STS", "BAR", "ACQBULK"]
cur, hist, order = None, {}, []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        k = re.search(r"(\w+_kernel)(<[^(]*>)?", d)
        cur = (k.group(1) + (k.group(2) or "")) if k else d[:60]
        cur = re.sub(r"__nv_bfloat16", "bf16", cur)
        if cur not in hist:
            hist[cur] = collections.Counter()
            order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        hist[cur][op] += 1
        hist[cur]["_total"] += 1
        if op in ("UTMALDG", "UTMASTG", "UTMAREDG", "ATOMG", "REDG", "RED"):
            hist[cur][op + m.group(2)] += 1
print("SASS mnemonic counts per kernel of myconvnet_b200/libmcn.so (sm_100a); '-' = none")
print("%-58s %6s " % ("kernel", "instr") + " ".join("%7s" % k[:7] for k in KEYS))
for name in order:
    h = hist[name]
    print("%-58s %6d " % (name[:58], h["_total"]) + " ".join("%7s" % (h[k] if h[k] else "-") for k in KEYS))
print()
print("TMA / atomic variants seen:")
var = collections.Counter()
for name in order:
    for k, v in hist[name].items():
        if "." in k:
            var[k] += v
for k, v in sorted(var.items()):
    print("  %-40s %d" % (k, v))
tot = collections.Counter()
for h in hist.values():
    tot.update({k: v for k, v in h.items() if "." not in k})
print()
print("library totals: " + ", ".join("%s=%d" % (k, tot[k]) for k in KEYS if tot[k]))
print("HMMA (legacy mma.sync) instructions: %d" % tot["HMMA"])
