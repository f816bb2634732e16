#!/usr/bin/env python
"""Hottest SASS instructions (by stall samples) of a `ncu --page source --csv` dump, in program order."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == 'Address')
iS = hdr.index('Source'); iSm = hdr.index('# Samples'); iI = hdr.index('Instructions Executed')
data = [r for r in rows if len(r) == len(hdr) and r[0] != 'Address']
tot = sum(int(r[iSm]) for r in data)
print('total samples', tot, 'static', len(data))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][iSm]))[:n]
for i in sorted(idx):
    r = data[i]
    print(f'{i:5d} {100*int(r[iSm])/tot:5.1f}% {r[iI]:>9s}  {r[iS][:100]}')
