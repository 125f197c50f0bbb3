#!/usr/bin/env python
"""Reduce an `ncu --metrics ... --csv` log of the staged kernels to one line per launch:
duration, DRAM bytes moved, DRAM throughput (% of peak), issue-slot and pipe utilisation."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
by = collections.OrderedDict()
for r in rows:
    by.setdefault((r[0], r[4].split("(")[0], r[8]), {})[r[12]] = float(r[14].replace(",", ""))
print(f"{'kernel':34s} {'grid':>12s} {'us':>9s} {'DRAM MB':>9s} {'DRAM GB/s':>10s} {'dram %':>7s} {'issue %':>8s} {'XU %':>6s} {'ALU %':>6s} {'FMA %':>6s} {'LSU %':>6s}")
for (i, name, grid), m in by.items():
    us = m.get("gpu__time_duration.sum", 0) / 1e3
    mb = (m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)) / 1e6
    print(f"{name[:34]:34s} {grid:>12s} {us:9.1f} {mb:9.1f} {mb / max(us, 1e-9) * 1e3:10.1f} "
          f"{m.get('dram__throughput.avg.pct_of_peak_sustained_elapsed', 0):7.1f} {m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):8.1f} "
          f"{m.get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 0):6.1f} {m.get('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 0):6.1f} "
          f"{m.get('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 0):6.1f} {m.get('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 0):6.1f}")
