"""Per-kernel time and DRAM traffic from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = collections.defaultdict(dict)
for r in csv.DictReader(lines):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    if r['Metric Name'] == 'gpu__time_duration.sum':
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    else:
        v = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}[u] * v
    rows[r['ID']][r['Metric Name']] = v
    rows[r['ID']]['name'] = re.sub(r'\(.*', '', r['Kernel Name'])
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in rows.values():
    a = agg[d['name']]
    a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0); a[2] += d.get('dram__bytes_read.sum', 0); a[3] += d.get('dram__bytes_write.sum', 0)
tot = sum(a[1] for a in agg.values())
print(f"launches {sum(a[0] for a in agg.values())}  total {tot:.0f} us")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 18]:
    print(f"{a[1]/tot*100:5.1f}% n={a[0]:3d} avg {a[1]/a[0]:7.1f} us  read {a[2]/a[0]:7.1f} MB  write {a[3]/a[0]:7.1f} MB  -> {(a[2]+a[3])/a[1]*1e3 if a[1] else 0:6.0f} GB/s  {k[:70]}")
