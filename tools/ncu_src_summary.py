"""Summarise an `ncu --page source --csv` export: stall reasons, opcode mix, hottest instructions."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx['# Samples']].isdigit()]
print(rows[0][1][:100])
tot = Counter()
ns = 0
for r in data:
    ns += int(r[idx['# Samples']])
    for s in stalls:
        if r[idx[s]].isdigit():
            tot[s] += int(r[idx[s]])
print('samples', ns, ' '.join('%s %.1f%%' % (s[6:], 100 * v / ns) for s, v in tot.most_common(9)))
c, ce = Counter(), Counter()
for r in data:
    t = r[idx['Source']].split()
    op = t[1] if t[0].startswith('@') else t[0]
    c[op] += int(r[idx['# Samples']])
    ce[op] += int(r[idx['Instructions Executed']])
te = sum(ce.values())
print('executed', te, ' '.join('%s %.1f%%(s%.1f%%)' % (op, 100 * v / te, 100 * c[op] / ns) for op, v in ce.most_common(14)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:n]:
    st = {s[6:]: int(r[idx[s]]) for s in stalls if r[idx[s]].isdigit() and int(r[idx[s]]) > 0}
    print(r[idx['# Samples']].rjust(6), r[idx['Source']][:64].ljust(64), dict(sorted(st.items(), key=lambda x: -x[1])[:2]))
