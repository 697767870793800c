"""Summarises an ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv): per kernel name the
launch count, total and average duration and share.  Usage: python profiles/launch_summary.py <csv> [skip_first_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[hdr + 1 + skip:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ki])[:80]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print("# %s: %d launches, %.3f ms in total (serialised, cold cache: compare SHARES)" % (sys.argv[1].split("/")[-1], sum(cnt.values()), T / 1e6))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:16]:
    print("%-82s n=%5d total %10.3f ms  avg %9.4f ms  %5.1f%%" % (k, cnt[k], v / 1e6, v / 1e6 / cnt[k], 100 * v / T))
