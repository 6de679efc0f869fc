"""Per-tile timeline of gemm_tma_kernel (consumer warp 0 of CTA 0) on the short-K shapes of the batched factorisation.
Needs a library built with -DFFGP_TG_TRACE:
  nvcc <flags of csrc/build.py> -DFFGP_TG_TRACE -o /tmp/libffgp_trace.so csrc/dense_gp.cu csrc/kron.cu ;  FFGP_LIB=/tmp/libffgp_trace.so python tools/tg_trace.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
L = ctypes.CDLL(os.environ.get('FFGP_LIB', 'gpurun_out/libffgp_trace.so'))
L.ffgp_gemm_f64.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int,
                            ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                            ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
g = torch.Generator(device='cuda').manual_seed(3)
st = torch.cuda.current_stream().cuda_stream
def run(name, akm, bkm, kmode, M, N, K, batch, lower=0, beta=0.0):
    A = torch.randn(batch, *((M, K) if akm else (K, M)), generator=g, dtype=torch.float64, device='cuda')
    Bm = torch.randn(batch, *((N, K) if bkm else (K, N)), generator=g, dtype=torch.float64, device='cuda')
    C = torch.zeros(batch, M, N, dtype=torch.float64, device='cuda')
    for _ in range(3):
        rc = L.ffgp_gemm_f64(akm, bkm, A.data_ptr(), A.shape[2], A.shape[1] * A.shape[2], Bm.data_ptr(), Bm.shape[2], Bm.shape[1] * Bm.shape[2],
                             C.data_ptr(), N, M * N, M, N, K, 1.0, beta, lower, kmode, batch, st)
        assert rc == 0
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (2 * 32 * 6))()
    assert L.ffgp_debug_tg_trace(buf) == 0
    print(f'{name}: per tile (clk): init | first stage wait | main loop | epilogue | next-tile wait | total      (warp 0, then warp 4 with its main-loop start relative to warp 0)')
    t0 = [[buf[n * 6 + s] for s in range(6)] for n in range(32)]
    print('   first tile, warp 0, clk per k-step:', [t0[24 + k + 1][0] - t0[24 + k][0] for k in range(7)])
    for wv in (0, 1):
        t = [[buf[(wv * 32 + n) * 6 + s] for s in range(6)] for n in range(32)]
        for n in range(1, 4):
            r = t[n]
            if r[5] <= r[0]: break
            print(f'   w{4 * wv} tile {n}: {r[1]-r[0]:6d} {r[2]-r[1]:6d} {r[3]-r[2]:7d} {r[4]-r[3]:6d} {r[5]-r[4]:6d} | {t[n+1][0]-r[0] if t[n+1][0] > r[0] else r[5]-r[0]:7d}   offset {r[2]-t0[n][2]:7d}')
nb = 8 * 148
run('dense 128-cube  A[i][p] B[j][p]', 1, 1, 0, 128, 128, 128, nb)
run('K_LE_COL 128 (L21 = A21 M11^T)', 1, 1, 2, 128, 128, 128, nb)
run('K_LE_ROW 128 (T = M22 L21)', 1, 0, 1, 128, 128, 128, nb)
run('K_GE_ROW 128 (S = M^T M block)', 0, 0, 4, 128, 128, 128, nb)
run('K_GE_COL 128 (M21 = -T M11)', 1, 0, 3, 128, 128, 128, nb)
run('dense 256-deep tiles', 1, 1, 0, 256, 256, 256, 2 * 148)
run('syrk lower 256 beta=1', 1, 1, 0, 256, 256, 256, 3 * 148, lower=1, beta=1.0)
run('S = M^T M 512 lower', 0, 0, 4, 512, 512, 512, 148, lower=1)
