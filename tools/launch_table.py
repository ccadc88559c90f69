"""Per-kernel totals of one steady-state step from an ncu launch list (gpu__time_duration.sum csv)."""
import csv, collections, re, sys
path = sys.argv[1]; step_idx = int(sys.argv[2]) if len(sys.argv) > 2 else -2
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if 'normalize_2nd' in r['Kernel Name']] + [len(rows)]
s, e = idx[step_idx], idx[step_idx + 1] if step_idx + 1 < len(idx) and step_idx != -1 else len(rows)
agg, tot = collections.OrderedDict(), 0.0
for r in rows[s:e]:
    k = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('shgan::', '')
    v = float(r['Metric Value']) / 1e6
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f'step launches {e - s}, serialized total {tot:.3f} ms')
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{v:8.3f} ms {100 * v / tot:5.1f}%  x{c:3d}  {k[:80]}')
if len(sys.argv) > 3:
    for r in rows[s:e]:
        k = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('shgan::', '')
        print(f"{float(r['Metric Value']) / 1e3:9.1f} us  grid {r['Grid Size']:>14}  {k[:50]}")
