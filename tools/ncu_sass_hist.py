"""Histogram of executed instructions by opcode and by code region from `ncu --page source --csv --print-source sass`.
Usage: python tools/ncu_sass_hist.py sass.csv [envs]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
envs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
ops, total = Counter(), 0
lines = []
for r in rows[2:]:
    try:
        n = int(r[iex])
    except Exception:
        continue
    src = r[isrc].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ops[op] += n
    total += n
    lines.append((n, src))
print(f"total executed warp-instructions: {total}  per env: {total/envs:.0f}")
for op, n in ops.most_common(40):
    print(f"{op:12s} {n:14d} {100*n/total:6.2f}%  per-env {n/envs:9.0f}")
# segment by execution count plateaus (loop nesting)
print("\n-- execution-count classes (count -> #static instrs, share)")
cls = Counter()
stat = Counter()
for n, _ in lines:
    cls[n] += n
    stat[n] += 1
for n, tot in sorted(cls.items(), key=lambda kv: -kv[1])[:12]:
    print(f"exec {n:12d} x {stat[n]:4d} static = {100*tot/total:6.2f}%")
