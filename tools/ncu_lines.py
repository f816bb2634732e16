#!/usr/bin/env python
"""Per-source-line dynamic instruction counts and stall samples of one kernel.
Joins `nvdisasm -g -c <cubin>` (line info; cubin from `cuobjdump -xelf all libnid_b200.so`) with
`ncu -i rep --page source --csv --kernel-name regex:<k>` by instruction order.
usage: ncu_lines.py <nvdisasm.txt> <mangled-kernel-name> <ncu_source.csv> [top_n] [pixels]"""
import collections, csv, re, sys

dis, kern, src = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
npx = float(sys.argv[5]) if len(sys.argv) > 5 else 0
lines = open(dis).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
cur = None
stack = []
sass = []  # (line_no, inline-chain, text)
for l in lines[start + 1:]:
    if l.startswith("//---") or l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inl" if "inlined at" in m.group(3) else "")
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        sass.append((cur, m.group(2)))
rows = list(csv.reader(open(src)))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
first = data[0][0]
for i in range(1, len(data)):
    if data[i][0] == first:
        data = data[:i]
        break
iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iT = hdr.index("Thread Instructions Executed")
assert len(data) == len(sass), (len(data), len(sass))
per = collections.OrderedDict()
for (loc, txt), r in zip(sass, data):
    d = per.setdefault(loc, [0, 0, 0, collections.Counter()])
    d[0] += int(r[iI]); d[1] += int(r[iS]); d[2] += int(r[iT])
    op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
    d[3][op.split(".")[0]] += int(r[iI])
ti = sum(d[0] for d in per.values()); ts = sum(d[1] for d in per.values()) or 1
print(f"{kern}: {ti} warp-instr, {ts} samples" + (f", {32 * ti / npx:.1f} lane-instr per pixel" if npx else ""))
srcs = {}
for loc, d in sorted(per.items(), key=lambda kv: -kv[1][0])[:topn]:
    fn, ln, inl = loc if loc else ("?", 0, "")
    if fn not in srcs:
        try:
            srcs[fn] = open("/root/repo/nid-pose-estimation_b200/csrc/" + fn).read().splitlines()
        except OSError:
            srcs[fn] = []
    text = srcs[fn][ln - 1].strip()[:70] if 0 < ln <= len(srcs[fn]) else ""
    mix = " ".join(f"{k}:{v * 100 // d[0]}" for k, v in d[3].most_common(3))
    print(f"{fn}:{ln:4d} {100 * d[0] / ti:5.1f}% instr {100 * d[1] / ts:5.1f}% stall  [{mix}]  {text}")
