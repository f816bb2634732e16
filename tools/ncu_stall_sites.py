#!/usr/bin/env python
"""Top sites of one stall reason in a `ncu --page source --csv` dump (first kernel instance), with opcode mix."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
reason = sys.argv[2] if len(sys.argv) > 2 else 'stall_long_sb'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
hdr = next(r for r in rows if r and r[0] == 'Address')
data = [r for r in rows if len(r) == len(hdr) and r[0] != 'Address']
# the dump concatenates the captured launches of the kernel: keep the first
first = data[0][0]
for i in range(1, len(data)):
    if data[i][0] == first:
        data = data[:i]
        break
iS = hdr.index('Source'); iL = hdr.index(reason); iI = hdr.index('Instructions Executed'); iA = hdr.index('# Samples')
tot = sum(int(r[iL]) for r in data); alls = sum(int(r[iA]) for r in data)
print(f'{reason}: {tot} of {alls} samples; static instructions {len(data)}; executed {sum(int(r[iI]) for r in data)}')
idx = sorted(range(len(data)), key=lambda i: -int(data[i][iL]))[:n]
for i in sorted(idx):
    r = data[i]
    print(f"{i:5d} {int(r[iL]):5d} {100*int(r[iL])/max(tot,1):5.1f}% {r[iI]:>8s} {r[iS][:90]}")
