#!/usr/bin/env python
"""Per-source-line view of an ncu report (needs -lineinfo + --import-source on):
     python tools/ncu_lines.py report.ncu-rep [top N] [ncu import filters, e.g. --launch-skip 1 --launch-count 1]
   prints executed warp instructions and stall samples per file:line, largest first, and per file."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
extra = sys.argv[3:]            # e.g. --launch-skip 1 --launch-count 1 to look at one kernel of the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + extra, capture_output=True, text=True).stdout
cur = None
hdr = None
lines = {}
stalls = defaultdict(lambda: defaultdict(int))
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Path":
        cur = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        continue
    if hdr is None or cur is None or not row[0].isdigit():
        continue
    ln = int(row[0])
    if len(row) != len(hdr):                 # unescaped quotes in the source text: the metric columns are still the last ones
        row = row[:2] + row[len(row) - (len(hdr) - 2):]
    try:
        inst = int(row[hdr.index("Instructions Executed")] or 0)
        smp = int(row[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    key = (cur, ln)
    a = lines.setdefault(key, [0, 0, row[1].strip()[:110]])
    a[0] += inst
    a[1] += smp
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and row[i] not in ("", "0"):
            stalls[key][h] += int(row[i])
tot_i = sum(v[0] for v in lines.values())
tot_s = sum(v[1] for v in lines.values())
print(f"total warp instructions {tot_i}, samples {tot_s}")
perfile = defaultdict(lambda: [0, 0])
for (f, ln), v in lines.items():
    perfile[f][0] += v[0]
    perfile[f][1] += v[1]
for f, v in sorted(perfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:28s} inst {100 * v[0] / tot_i:5.1f} %   samples {100 * v[1] / max(tot_s, 1):5.1f} %")
print("\n-- by samples")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ", ".join(f"{k[6:]} {n}" for k, n in sorted(stalls[(f, ln)].items(), key=lambda kv: -kv[1])[:3])
    print(f"{f}:{ln:<5d} inst {100 * v[0] / tot_i:5.2f} %  smp {100 * v[1] / max(tot_s, 1):5.2f} %  [{st}]  {v[2]}")
print("\n-- by instructions")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:<5d} inst {100 * v[0] / tot_i:5.2f} %  smp {100 * v[1] / max(tot_s, 1):5.2f} %  {v[2]}")
