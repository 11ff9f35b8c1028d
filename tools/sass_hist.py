"""Summarise an `ncu --page source --csv` dump: executed-instruction histogram by opcode and the hottest SASS lines.
usage: python tools/sass_hist.py file.csv [n_hot_lines]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
h = rows[hi]
ie, si, ss = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
ops, samp = collections.Counter(), collections.Counter()
tot = totS = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    n, s, src = int(r[ie]), int(r[ss]), r[si].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else src
    ops[op] += n
    samp[op] += s
    tot += n
    totS += s
    lines.append((n, s, src))
print("total warp instructions", tot, "stall samples", totS)
for op, n in ops.most_common(22):
    print(f"{op:12s} {n:12d} {100 * n / tot:5.1f}%   samples {100 * samp[op] / max(1, totS):5.1f}%")
if nhot:
    print("--- hottest lines by samples")
    for n, s, src in sorted(lines, key=lambda x: -x[1])[:nhot]:
        print(f"{s:7d} samples {n:10d} exec  {src}")
