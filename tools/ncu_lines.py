"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:KERNEL` dump per CUDA source
line: executed warp instructions and stall samples, sorted by line, with the share of the kernel total."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hi = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Line No"][0]
hdr = rows[hi]
ie = hdr.index("Instructions Executed")
ss = hdr.index("Warp Stall Sampling (All Samples)")
lines = []
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[0].strip().isdigit():
        continue
    n = int(r[ie]) if r[ie].isdigit() else 0
    s = int(r[ss]) if r[ss].isdigit() else 0
    lines.append((int(r[0]), n, s, r[1].strip()))
total = sum(n for _, n, _, _ in lines) or 1
stot = sum(s for _, _, s, _ in lines) or 1
print(f"total warp instructions {total}, stall samples {stot}")
for ln, n, s, src in sorted(lines, key=lambda t: -t[1])[:top]:
    print(f"{ln:5d} {n:11d} {n / total:6.1%}  samples {s / stot:6.1%}  {src[:110]}")
