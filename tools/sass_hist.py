"""Opcode histogram (executed instructions, stall samples) of one kernel from an ncu source-page CSV.
usage: ncu -i x.ncu-rep --page source --csv --print-source sass --kernel-name regex:K > k.csv; python tools/sass_hist.py k.csv [top-lines]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == 'Kernel Name':
        break
    data.append(r)
ia = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
tot_s = sum(int(r[isamp]) for r in data) or 1; tot_e = sum(int(r[iex]) for r in data) or 1
print('total samples', tot_s, 'warp-inst', tot_e, 'sass lines', len(data))
c = Counter(); s = Counter()
for r in data:
    t = r[ia].split()
    op = t[1] if t[0].startswith('@') else t[0]
    op = op.split('.')[0]
    c[op] += int(r[iex]); s[op] += int(r[isamp])
for op, v in c.most_common(22):
    print(f'{op:12s} exec {v:>12d} {100*v/tot_e:5.1f}%   samples {100*s[op]/tot_s:5.1f}%')
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if n:
    print('--- hottest SASS lines by samples')
    order = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:n]
    for i in sorted(order):
        print(f'{i:5d} {int(data[i][isamp]):6d} {int(data[i][iex]):>10d}  {data[i][ia].strip()[:110]}')
