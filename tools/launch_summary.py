"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and
share per kernel.

    python tools/launch_summary.py <launches.csv> [title]
"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) != len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void (pdenet::)?", "", r[idx["Kernel Name"]]).split("(")[0].replace("(int)", "")
    name = name.replace("pdenet::", "")
    v = float(r[idx["Metric Value"]].replace(",", ""))
    unit = r[idx["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
              "second": 1e3}[unit]
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2] + "\n")
print("| kernel | launches | total ms | share under ncu |\n|---|---|---|---|")
for k in sorted(tot, key=lambda k: -tot[k]):
    print("| `%s` | %d | %.1f | %.1f%% |" % (k[:110], cnt[k], tot[k], 100 * tot[k] / total))
print("\n%d launches, %.1f ms in total." % (sum(cnt.values()), total))
