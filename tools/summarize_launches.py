"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time and count per kernel, optionally
only the last `--evals`-th share of the launches (the warm evaluation)."""
import argparse, csv, collections, re, sys
ap = argparse.ArgumentParser(); ap.add_argument('csv'); ap.add_argument('--last-frac', type=float, default=1.0)
ap.add_argument('--by-grid', action='store_true')
a = ap.parse_args()
rows = []
with open(a.csv) as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.DictReader(lines)
for row in r:
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    ns = v * {'ns': 1, 'us': 1e3, 'usecond': 1e3, 'nsecond': 1, 'ms': 1e6, 'msecond': 1e6, 'second': 1e9}[u]
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    if a.by_grid:
        name += ' grid=' + row['Grid Size']
    rows.append((name, ns))
n0 = int(len(rows) * (1 - a.last_frac))
rows = rows[n0:]
tot = sum(ns for _, ns in rows)
agg = collections.OrderedDict()
for name, ns in rows:
    c = agg.setdefault(name, [0, 0.0]); c[0] += 1; c[1] += ns
print(f'{len(rows)} launches, total {tot / 1e6:.3f} ms')
for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{ns / 1e6:9.3f} ms {100 * ns / tot:5.1f}%  x{cnt:<5d} {name[:150]}')
