"""Per-launch table from an ncu csv with several metrics (time, DRAM bytes, pipe utilisation):
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --csv --log-file f.csv <cmd>
  python tools/summarize_metrics.py f.csv [--skip N] [--agg]"""
import argparse, csv, collections, re
ap = argparse.ArgumentParser(); ap.add_argument('csv'); ap.add_argument('--skip', type=int, default=0)
ap.add_argument('--agg', action='store_true'); ap.add_argument('--filter', default='')
a = ap.parse_args()
with open(a.csv) as f:
    lines = [l for l in f if not l.startswith('==')]
UNIT = {'ns': 1e-9, 'nsecond': 1e-9, 'us': 1e-6, 'usecond': 1e-6, 'ms': 1e-3, 'msecond': 1e-3, 'second': 1, 's': 1,
        'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, '%': 1, '': 1}
launches = collections.OrderedDict()
for row in csv.DictReader(lines):
    i = int(row['ID'])
    L = launches.setdefault(i, {'name': re.sub(r'^void ', '', re.sub(r'\(.*', '', row['Kernel Name'])), 'grid': row['Grid Size']})
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    L[row['Metric Name']] = v * UNIT.get(row['Metric Unit'], 1)
rows = [L for i, L in launches.items() if i >= a.skip and a.filter in L['name']]
def fmt(L):
    t = L.get('gpu__time_duration.sum', 0.0)
    rd, wr = L.get('dram__bytes_read.sum', 0.0), L.get('dram__bytes_write.sum', 0.0)
    extra = ' '.join(f'{k.split(".")[0].replace("sm__", "").replace("_cycles_active", "")}={v:5.1f}' for k, v in L.items()
                     if k not in ('name', 'grid', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'count'))
    return (f"{t * 1e6:9.1f} us  rd {rd / 1e6:9.2f} MB  wr {wr / 1e6:9.2f} MB  {(rd + wr) / max(t, 1e-12) / 1e9:7.0f} GB/s  {extra}  "
            f"{L.get('count', '')} {L['name'][:70]} {L['grid']}")
if a.agg:
    agg = collections.OrderedDict()
    for L in rows:
        A = agg.setdefault(L['name'], {'name': L['name'], 'grid': '', 'count': 0})
        A['count'] += 1
        for k, v in L.items():
            if k in ('name', 'grid', 'count'):
                continue
            if 'pct' in k:
                A[k] = A.get(k, 0.0) + v * L.get('gpu__time_duration.sum', 0.0)      # time-weighted
            else:
                A[k] = A.get(k, 0.0) + v
    for A in agg.values():
        for k in list(A):
            if 'pct' in k:
                A[k] /= max(A.get('gpu__time_duration.sum', 0.0), 1e-12)
        A['count'] = f"x{A['count']}"
    rows = sorted(agg.values(), key=lambda L: -L.get('gpu__time_duration.sum', 0.0))
tot = sum(L.get('gpu__time_duration.sum', 0.0) for L in rows)
trd = sum(L.get('dram__bytes_read.sum', 0.0) + L.get('dram__bytes_write.sum', 0.0) for L in rows)
print(f'{len(rows)} rows, total {tot * 1e3:.3f} ms, DRAM traffic {trd / 1e9:.3f} GB')
for L in rows:
    print(fmt(L))
