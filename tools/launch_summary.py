"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
seq = []
for row in r:
    if len(row) <= vi:
        continue
    v = float(row[vi].replace(',', ''))
    v = v / 1e3 if row[ui] == 'ns' else (v * 1e3 if row[ui] == 'ms' else v)
    seq.append((re.sub(r'\(.*', '', row[ki])[:56], v, row[gi]))
tot = sum(v for _, v, _ in seq)
print(f"{len(seq)} launches, {tot:.1f} us of kernel time (cold-cache, serialised under ncu: compare SHARES)")
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v, _ in seq:
    agg[k][0] += 1
    agg[k][1] += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}% {n:4d}x  {k}")
if len(sys.argv) > 3:
    print([(round(v, 1), g) for k, v, g in seq if sys.argv[3] in k])
