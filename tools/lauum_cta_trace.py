"""Per-CTA durations of the batched S = M^T M launch (one tile per CTA), grouped by tile (ti, tj).  Needs -DFFGP_TG_TRACE:
   FFGP_LIB=tools/libffgp_trace.so python tools/lauum_cta_trace.py"""
import ctypes, os, sys, collections
import torch
L = ctypes.CDLL(os.environ.get('FFGP_LIB', 'tools/libffgp_trace.so'))
L.ffgp_gemm_f64.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int,
                            ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                            ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
batch, n = 1024, 512
M = torch.tril(torch.randn(batch, n, n, dtype=torch.float64, device='cuda')); S = torch.zeros_like(M)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    assert L.ffgp_gemm_f64(0, 0, M.data_ptr(), n, n * n, M.data_ptr(), n, n * n, S.data_ptr(), n, n * n, n, n, n, 1.0, 0.0, 1, 4, batch, st) == 0
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (4096 * 4))()
assert L.ffgp_debug_tg_cta(buf) == 0
names = [(i, j) for i in range(4) for j in range(i + 1)]
acc = collections.defaultdict(list)
for c in range(1480, 4096):          # skip the first wave (cold start)
    t = [buf[c * 4 + s] for s in range(4)]
    acc[names[c % 10]].append((t[1] - t[0], t[2] - t[1], t[3] - t[2], t[3] - t[0]))
print('tile   k-blocks   init+first stage | main loop | epilogue | total   (median clk over CTAs; CTA launch + barrier init + exit not included)')
tot = 0
for (i, j), v in sorted(acc.items()):
    med = [sorted(x[k] for x in v)[len(v) // 2] for k in range(4)]
    tot += med[3]
    print(f'({i},{j})   {4 - i}         {med[0]:7d} {med[1]:9d} {med[2]:8d} {med[3]:8d}')
print('sum over the 10 tiles of a problem:', tot, 'clk')
