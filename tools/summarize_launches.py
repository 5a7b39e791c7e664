"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:64]:64s} n={cnt[k]:4d} total={v / 1e3:10.1f}us avg={v / cnt[k] / 1e3:9.1f}us share={v / s:6.1%}")
