"""Summarise a FFGP_TRACE timeline (tools/timeline_c2.py): per right-looking step the completion times of the column
update, the two diagonal-block kernels, the TRSM and the trailing update, then the background / post-factorisation
launches."""
import sys, collections
rows = []
for l in open(sys.argv[1]):
    if not l.startswith('TRACE'):
        continue
    _, ms, st, what, a, b = l.split()
    rows.append((float(ms), st, what, int(a), int(b)))
ev = collections.defaultdict(dict)
k = -1
for ms, st, what, a, b in rows:
    if what == 'a:colupd': k += 1; ev[k]['a'] = ms
    elif what == 'b:trsm': ev[k]['trsm'] = ms
    elif what == 'c:syrk': ev[k]['c'] = ms
    elif what == 'base' and k >= 0: ev[k].setdefault('base', []).append(ms)
print('step  colupd_end  base_ends            trsm_end  syrk_end  step_len(ms)')
for k in sorted(ev):
    e = ev[k]
    print(f"{k:3d}  {e.get('a', 0):9.3f}  {' '.join('%.3f' % x for x in e.get('base', [])):20s} {e.get('trsm', 0):8.3f}  {e.get('c', 0):8.3f}  "
          f"{(e.get('a', 0) - ev[k - 1].get('a', 0)) if k > 0 else 0:6.3f}")
print('background / post launches (completion ms, label, M, N*1e5+K):')
for r in rows:
    if r[2].startswith('bg') or r[2].startswith('post') or r[2] == 'gemm':
        print(f'  {r[0]:9.3f} {r[2]:12s} {r[3]:6d} {r[4]}')
