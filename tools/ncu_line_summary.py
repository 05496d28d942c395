"""Aggregate ncu per-instruction stall samples by CUDA source line.

    python tools/ncu_line_summary.py <src.csv from `ncu --page source --csv`> <cubin> <mangled-substring> [n]

Maps SASS offsets to source lines with `nvdisasm -g` (needs -lineinfo)."""
import csv
import re
import subprocess
import sys
from collections import Counter

src_csv, cubin, sym = sys.argv[1:4]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 30
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if re.match(r"\s*\.section\s+\.text\.", l) and sym in l)
cur, a2l = None, {}
for l in dis[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
    m2 = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m2:
        a2l[int(m2.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
base = int(data[0][idx["Address"]], 16)
smp, ex = Counter(), Counter()
seen = set()
for r in data:
    off = int(r[idx["Address"]], 16) - base
    if off in seen:
        continue  # the export repeats each kernel's rows
    seen.add(off)
    ln = a2l.get(off)
    smp[ln] += int(r[idx["# Samples"]])
    ex[ln] += int(r[idx["Instructions Executed"]])
tot, tex = sum(smp.values()), sum(ex.values())
cache = {}
for ln, v in smp.most_common(n):
    text = ""
    if ln:
        if ln[0] not in cache:
            try:
                cache[ln[0]] = open(ln[0]).read().split("\n")
            except OSError:
                cache[ln[0]] = []
        if ln[1] - 1 < len(cache[ln[0]]):
            text = cache[ln[0]][ln[1] - 1].strip()[:90]
    print("%5.1f%% smp %5.1f%% ex  %s:%s  %s" % (100 * v / tot, 100 * ex[ln] / tex, ln[0].split("/")[-1] if ln else None, ln[1] if ln else "", text))
