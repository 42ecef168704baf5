#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py REPORT KERNEL_REGEX [N]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
ix = {h: i for i, h in enumerate(rows[0])}
r = rows[2]
for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active"):
    if k in ix: print(k, r[ix[k]], rows[1][ix[k]])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = [r for r in rows if len(r) > 5 and r[0] == 'Line No'][0]; nH = len(hdr)
cur = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0].isdigit():
        extra = len(r) - nH
        vals = r[:1] + [','.join(r[1:2 + extra])] + r[2 + extra:]
        agg.append((cur, int(r[0]), int(vals[6]), int(vals[7]), vals[1]))
ts = sum(a[2] for a in agg); ti = sum(a[3] for a in agg)
print("total samples", ts, "inst", ti)
print("--- by samples")
for f, ln, s, i, t in sorted(agg, key=lambda x: -x[2])[:N]:
    print(f'{f[:16]:16s}:{ln:4d} samp {100*s/ts:5.1f}% inst {100*i/ti:5.1f}%  {t[:100]}')
print("--- by instructions")
for f, ln, s, i, t in sorted(agg, key=lambda x: -x[3])[:N]:
    print(f'{f[:16]:16s}:{ln:4d} inst {100*i/ti:5.1f}% samp {100*s/ts:5.1f}%  {t[:100]}')
