"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    k = re.sub(r"void |fidm::|\(.*", "", row["Kernel Name"])
    tot[k] += v
    cnt[k] += 1
T = sum(tot.values())
print(f"total {T:.1f} us over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:10.1f} us {100 * v / T:5.1f}% n={cnt[k]:4d}  {k[:90]}")
