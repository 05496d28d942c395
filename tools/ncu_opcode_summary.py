"""python tools/ncu_opcode_summary.py <source-page.csv>
Aggregate an `ncu --page source --csv` export by SASS opcode: executed warp instructions and the
sampled stall reasons attributed to each opcode."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
names = ['stall_barrier', 'stall_branch_resolving', 'stall_dispatch', 'stall_long_sb', 'stall_math', 'stall_mio',
         'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_wait', 'stall_lg']
tot, cnt, stalls, allinst = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter), 0
for r in rows[2:]:
    if len(r) < len(hdr) or r[ix['Source']] == 'Source':
        continue
    src = r[ix['Source']].strip()
    toks = src.split()
    op = toks[0] if toks else '?'
    if op.startswith('@') and len(toks) > 1:
        op = toks[1]
    op = op.split('.')[0]
    n = int((r[ix['Instructions Executed']] or '0').replace(',', ''))
    cnt[op] += n
    allinst += n
    for s in names:
        v = int((r[ix[s]] or '0').replace(',', ''))
        stalls[op][s] += v
        tot[s] += v
print('total warp instructions', allinst)
for op, n in cnt.most_common(16):
    print('%-8s %12d %5.1f%%  %s' % (op, n, 100 * n / allinst,
                                    {s.replace('stall_', ''): v for s, v in stalls[op].items() if v > 200}))
print({s.replace('stall_', ''): v for s, v in tot.items()})
