#!/usr/bin/env python
"""Instruction mix / stall samples per SASS opcode from `ncu --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kern = None
blocks = collections.OrderedDict()
hdr = None
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]; blocks[kern] = []; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        blocks[kern].append(r)
for kern, data in blocks.items():
    iS = hdr.index('Source'); iI = hdr.index('Instructions Executed'); iSm = hdr.index('# Samples')
    tot = sum(int(r[iI]) for r in data)
    print(f"== {kern}: static {len(data)} SASS, dynamic {tot} warp-instr")
    ops = collections.Counter(); samp = collections.Counter()
    for r in data:
        toks = r[iS].split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        op = '.'.join(op.split('.')[:2]) if op.startswith(('I2F', 'F2I', 'LDG', 'MUFU', 'LDS', 'STS', 'F2F')) else op.split('.')[0]
        ops[op] += int(r[iI]); samp[op] += int(r[iSm])
    ts = sum(samp.values()) or 1
    for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
        print(f"  {op:14s} {c:11d} {100*c/tot:5.1f}%   stall samples {100*samp[op]/ts:5.1f}%")
