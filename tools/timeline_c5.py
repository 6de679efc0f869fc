"""Per-launch timeline of one warm batched evaluation (BASELINE config 5), FFGP_TRACE=1: time between consecutive
launch completions on the stream, summed per kernel label.   FFGP_TRACE=1 python tools/timeline_c5.py [--batch 4096]"""
import argparse, collections, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get('TL_CHILD'):
    sys.path.insert(0, ROOT)
    import torch
    from fidelityfusion_b200 import _lib
    from fidelityfusion_b200.batched import batched_cigp_eval
    B = int(os.environ['TL_CHILD']); n, d, ns = 512, 8, 64
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64); w = torch.randn(B, d, 1, generator=g, dtype=torch.float64)
    y = torch.sin(3 * sum(x[..., k:k + 1] * w[:, k:k + 1, :] for k in range(d))) + 0.05 * torch.randn(B, n, 1, generator=g, dtype=torch.float64)
    ls = torch.exp(torch.rand(B, d, generator=g, dtype=torch.float64) * 2 - 1); sv = torch.ones(B, dtype=torch.float64)
    lb = torch.rand(B, generator=g, dtype=torch.float64) * 3; xs = torch.rand(B, ns, d, generator=g, dtype=torch.float64)
    x, y, ls, sv, lb, xs = (t.cuda() for t in (x, y, ls, sv, lb, xs))
    L = _lib.lib()
    for i in range(3):
        batched_cigp_eval(x, y, ls, sv, lb, xs); torch.cuda.synchronize()
        if i < 2:
            devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1); L.ffgp_trace_dump(); os.dup2(saved, 1); os.close(devnull)
        else:
            sys.stdout.flush(); L.ffgp_trace_dump()
    sys.exit(0)
ap = argparse.ArgumentParser(); ap.add_argument('--batch', type=int, default=4096); a = ap.parse_args()
r = subprocess.run([sys.executable, __file__], env=dict(os.environ, FFGP_TRACE='1', TL_CHILD=str(a.batch)), capture_output=True, text=True)
rows = [l.split() for l in r.stdout.splitlines() if l.startswith('TRACE')]
if not rows: print(r.stdout[-2000:], r.stderr[-2000:]); sys.exit(1)
tot = collections.OrderedDict(); prev = 0.0
print('  ms(end)   dt(ms)  label            M      N*1e5+K')
for _, ms, st, what, aa, bb in rows:
    ms = float(ms); key = f'{what} {aa} {bb}'
    print(f'{ms:9.3f} {ms - prev:8.3f}  {what:14s} {aa:>6s} {bb}')
    tot[what] = tot.get(what, 0.0) + ms - prev; prev = ms
print('--- per label (ms, % of the traced span; the first entry includes everything before the first traced launch)')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]): print(f'{k:16s} {v:8.3f}  {100 * v / prev:5.1f} %')
