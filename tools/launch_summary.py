"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv"""
import collections
import csv
import sys

for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[0] == 'ID':
            hdr = r
            continue
        if hdr is None:
            continue
        try:
            name = r[hdr.index('Kernel Name')]
            v = float(r[-1].replace(',', ''))
        except Exception:
            continue
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{f}: {tot / 1000:.1f} us in {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[1] / 1000:10.1f} us {v[0]:5d} {100 * v[1] / tot:5.1f}% {k[:100]}")
