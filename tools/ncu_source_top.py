#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples, and a per-region roll-up, from an .ncu-rep (source page)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
# first kernel only
start = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
blk = lines[start[0] + 1: start[1] if len(start) > 1 else None]
rows = list(csv.reader(io.StringIO("\n".join(blk))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[1:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, " instructions", len(data))
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print("stall totals:", {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
ranked = sorted(enumerate(data), key=lambda t: -int(t[1][ix["# Samples"]]))[:topn]
for n, r in sorted(ranked):
    st = {h[6:]: int(r[ix[h]] or 0) for h in stall_cols if int(r[ix[h]] or 0)}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{n:5d} {int(r[ix['# Samples']]):6d} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:70]:70s} {top}")
