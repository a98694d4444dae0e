"""Executed-instruction histogram by opcode from an `ncu --page source --csv` dump.  usage: ncu_ophist.py src.csv n_units"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]
ia, ie = hdr.index('Source'), hdr.index('Instructions Executed')
c = Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    e = int(r[ie])
    toks = r[ia].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    c[op.split('.')[0]] += e
    tot += e
print('total', tot, 'per unit', tot / units)
for op, v in c.most_common(30):
    print(f'{op:10s} {v / units:8.2f}')
