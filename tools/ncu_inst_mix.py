#!/usr/bin/env python
"""Executed warp-instruction mix of the first kernel of an .ncu-rep (source page): by opcode and by SASS region."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
blk = lines[start[0] + 1: start[1] if len(start) > 1 else None]
rows = list(csv.reader(io.StringIO("\n".join(blk))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[1:]
by_op = collections.Counter()
total = 0
for r in data:
    n = int(r[ix["Instructions Executed"]] or 0)
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2])
    by_op[op] += n
    total += n
print("total warp instructions", total)
for op, n in by_op.most_common(40):
    print(f"{op:28s} {n:12d} {100.0 * n / total:6.2f}%")
# regions of 100 SASS lines
print("-- per 100-line SASS region --")
for a in range(0, len(data), 100):
    n = sum(int(r[ix["Instructions Executed"]] or 0) for r in data[a:a + 100])
    if n:
        print(f"{a:5d}-{a + 99:5d} {n:12d} {100.0 * n / total:6.2f}%")
