#!/usr/bin/env python3
"""Instruction mix of a kernel from an `ncu --page source --csv --print-source sass` export (optionally .gz):
warp-level instructions executed per SASS opcode, with the share of the total.
usage: sass_mix.py source_sass.csv[.gz] [top]"""
import collections, csv, gzip, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = csv.reader(l for l in f if not l.startswith("=="))
hdr = None
mix = collections.Counter()
kernels = []
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels.append(r[1][:60])
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or len(kernels) > 1:
        continue
    src = r[hdr.index("Source")].strip()
    try:
        n = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    toks = src.split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    if not toks:
        continue
    mix[toks[0].rstrip(";")] += n
tot = sum(mix.values())
print("kernel: %s   (first launch of the report)" % (kernels[0] if kernels else "?"))
print("total warp-level instructions executed: %d" % tot)
for op, n in mix.most_common(top):
    print("%-28s %14d  %5.1f %%" % (op, n, 100.0 * n / tot))
