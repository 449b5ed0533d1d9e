"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share / count / min / median / max per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
d = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    if unit == 'ns':
        v /= 1e3
    elif unit == 'ms':
        v *= 1e3
    d[name].append(round(v, 1))
tot = sum(sum(v) for v in d.values())
print(f'launches {sum(len(v) for v in d.values())}  total {tot:.1f} us (cold-cache, serialised: compare shares)')
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))[:top]:
    v = sorted(v)
    print(f"{sum(v)/tot*100:5.1f}% n={len(v):3d} min {v[0]:6.1f} med {v[len(v)//2]:6.1f} max {v[-1]:6.1f} us  {k[:100]}")
