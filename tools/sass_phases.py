"""Split an `ncu --page source --csv` (SASS view) dump at BAR.SYNC instructions and report executed warp
instructions per phase with an opcode histogram -- a cheap per-phase instruction profile of a fused kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if len(r) > 1 and r[1] == "Source"][0]
hdr = rows[hi]
ic, sc = hdr.index("Instructions Executed"), hdr.index("Source")
data = [(r[sc].strip(), int(r[ic]) if r[ic].isdigit() else 0) for r in rows[hi + 1:] if len(r) > ic]
total = sum(n for _, n in data)
print("total warp instructions", total)
phase, acc, hist = 0, 0, collections.Counter()
def flush():
    global acc, hist
    top = ", ".join(f"{k}:{v / max(acc, 1):.0%}" for k, v in hist.most_common(9))
    print(f"phase {phase:2d}: {acc:10d} ({acc / total:6.1%})  {top}")
    acc, hist = 0, collections.Counter()
for src, n in data:
    op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0] if src else "?"
    op = op.split(".")[0]
    acc += n
    hist[op] += n
    if src.startswith("BAR.SYNC") or " BAR.SYNC" in src:
        flush()
        phase += 1
flush()
allh = collections.Counter()
for src, n in data:
    op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0] if src else "?"
    allh[op.split(".")[0]] += n
print("overall:", ", ".join(f"{k}:{v / total:.1%}" for k, v in allh.most_common(20)))
