"""Shared-memory wavefronts per CUDA source line from an `ncu --page source --csv --print-source cuda,sass` dump:
python tools/ncu_smem_lines.py dump.csv [top]  -- wavefronts, ideal wavefronts, excess (bank conflicts), instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = {}
fname = None
for i, r in enumerate(rows):
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > 3 and r[0] == "Line No":
        hdr = r
        iw, ii, ie, ix = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("Instructions Executed")
        if hdr[2] == "Address":  # the SASS view repeats the counters per instruction: use the CUDA view only
            pass
        for q in rows[i + 1:]:
            if len(q) <= iw or not q[0].strip().isdigit():
                if len(q) >= 1 and q[0] in ("File Path", "Function Name", "Line No"):
                    break
                continue
            key = (fname, int(q[0]))
            def num(s):
                try:
                    return int(float(s))
                except Exception:
                    return 0
            a = agg.setdefault(key, [0, 0, 0, 0, q[1].strip()[:90]])
            a[0] += num(q[iw]); a[1] += num(q[ii]); a[2] += num(q[ie]); a[3] += num(q[ix])
tot = [sum(a[k] for a in agg.values()) for k in range(4)]
print(f"shared wavefronts {tot[0]}, ideal {tot[1]}, excessive {tot[2]}, instructions {tot[3]}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:5d} wavefronts {a[0]:10d} ({a[0] / max(tot[0], 1):5.1%}) ideal {a[1]:10d} excess {a[2]:10d}  instr {a[3]:9d}  {a[4]}")
