"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    name = row['Kernel Name'].split('(')[0].replace('void ', '')
    if 'at::' in name:
        name = 'torch::' + name.split('::')[-1][:40]
    t = float(row['Metric Value']) / 1e6
    agg[name][0] += 1
    agg[name][1] += t
    tot += t
print(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t:10.3f} ms {100*t/tot:5.1f}%  x{c:4d}  {n}")
