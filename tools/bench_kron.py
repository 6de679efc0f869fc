"""HBM-roofline measurements of the Kronecker / coupling half of the hot path:
mode products at the C4 shapes (128 x 32 x 32 x 16), the fused core stage, the per-mode Jacobi eigensolver, the
whole HOGP loss + gradient, and the C3 coupling (Tensor_linear 1024 -> 4096) + cigp NLL with D = 4096 columns.
CUDA events, warm, L2 flushed between repetitions.   python tools/bench_kron.py [--json out.json]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument('--json', default='')
ap.add_argument('--reps', type=int, default=10)
ap.add_argument('--only', default='')
a = ap.parse_args()
torch.set_default_dtype(torch.float64)
from fidelityfusion_b200 import tensorly_compat as tl
from fidelityfusion_b200.MFGP_ver2023May import HOGP
from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
from fidelityfusion_b200.GaussianProcess.gp_computation_pack import Tensor_linear

HBM = 6457.7
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
flush_buf = torch.zeros(512 << 20, dtype=torch.uint8, device='cuda').view(torch.int64)


def timeit(fn, reps=a.reps, flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            flush_buf.sum()          # READ 512 MB: evicts the 126 MB L2 and leaves only clean lines behind (a write-flush
                                     # leaves dirty lines whose write-back would be charged to the kernel under test)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


rows = []


def report(name, sec, bytes_alg=None, flops=None):
    r = {'name': name, 'us': sec * 1e6}
    if bytes_alg:
        r['GBps'] = bytes_alg / sec * 1e-9
        r['hbm_frac'] = r['GBps'] / HBM
    if flops:
        r['TFLOPs'] = flops / sec * 1e-12
    rows.append(r)
    print(f"{name:58s} {sec * 1e6:10.1f} us" + (f"  {r['GBps']:8.1f} GB/s ({100 * r['hbm_frac']:5.1f}% of HBM)" if bytes_alg else '')
          + (f"  {r['TFLOPs']:6.2f} TFLOP/s" if flops else ''), flush=True)


g = torch.Generator().manual_seed(4)
shape = (128, 32, 32, 16)
T = torch.randn(*shape, generator=g).cuda()
numel = T.numel()
# calibration: what a 16 MiB -> 16 MiB streaming kernel can reach at THIS size (launch + ramp dominate 5 us of transfer)
report('[calibration] empty timed region', timeit(lambda: None))
report('[calibration] torch clone of the C4 tensor (16.8 MB in, 16.8 MB out)', timeit(lambda: T.clone()), 16 * numel)
Tbig = torch.randn(1024, 32, 32, 16, generator=g).cuda()
report('[calibration] torch clone, 8x larger tensor (134 MB in, 134 MB out)', timeit(lambda: Tbig.clone()), 16 * Tbig.numel())
if not a.only or 'mode' in a.only:
    for mode in (1, 2, 3):
        I = Tbig.shape[mode]
        U = torch.randn(I, I, generator=g).cuda()
        sec = timeit(lambda: tl._mode_dot_raw(Tbig, U, mode, False))
        report(f'mode_dot 1024x32x32x16 mode {mode} ({I}x{I})', sec, 16 * Tbig.numel() + 8 * I * I, 2 * Tbig.numel() * I)
        sec = timeit(lambda: tl._mode_gram_raw(Tbig, Tbig, mode))
        report(f'mode_gram 1024x32x32x16 mode {mode} ({I}x{I})', sec, 16 * Tbig.numel(), 2 * Tbig.numel() * I)
if not a.only or 'mode' in a.only:
    for mode, I in enumerate(shape):
        U = torch.randn(I, I, generator=g).cuda()
        sec = timeit(lambda: tl._mode_dot_raw(T, U, mode, False))
        report(f'mode_dot C4 mode {mode} ({I}x{I})', sec, 16 * numel + 8 * I * I, 2 * numel * I)
        sec = timeit(lambda: tl._mode_dot_raw(T, U, mode, True))
        report(f'mode_dot C4 mode {mode} ({I}x{I}) transposed factor', sec, 16 * numel + 8 * I * I, 2 * numel * I)
    for mode, I in enumerate(shape):
        sec = timeit(lambda: tl._mode_gram_raw(T, T, mode))
        report(f'mode_gram C4 mode {mode} ({I}x{I})', sec, 16 * numel, 2 * numel * I)
    # torch baseline (cuBLAS through tensordot + permute copies) for the same contraction
    for mode, I in enumerate(shape):
        U = torch.randn(I, I, generator=g).cuda()
        sec = timeit(lambda: torch.movedim(torch.tensordot(U, T, ([1], [mode])), 0, mode).contiguous())
        report(f'[torch tensordot+permute] mode {mode}', sec, 16 * numel + 8 * I * I, 2 * numel * I)

if not a.only or 'core' in a.only:
    sizes = list(shape)
    lam = torch.rand(sum(sizes), generator=g).cuda() + 0.1
    tau = torch.tensor([0.5]).cuda()
    sec = timeit(lambda: tl._kron_core(T, lam, sizes, tau, 0.0))
    report('kron_core (A, T1/A, 4 sums): read 8B, write 16B / element', sec, 24 * numel)
    sec = timeit(lambda: tl._kron_scale(T, lam, sizes, 1, 0, tau, 0.0, T.device))
    report('kron_scale: read 8B, write 8B / element', sec, 16 * numel)

if not a.only or 'eigh' in a.only:
    for n in (16, 32, 64, 128):
        x = torch.randn(n, 5, generator=g).cuda()
        K = torch.exp(-0.5 * torch.cdist(x, x) ** 2)
        sec = timeit(lambda: tl.eigh(K), flush=False)
        report(f'syevj n={n} (incl. python wrapper + info readback)', sec)
        sec = timeit(lambda: tl._eigh_launch(K), flush=False)
        report(f'syevj n={n} (kernel enqueue only, no status readback)', sec)
        sec = timeit(lambda: torch.linalg.eigh(K), flush=False)
        report(f'[torch.linalg.eigh / cuSOLVER] n={n}', sec)

if not a.only or 'hogp' in a.only:
    x = torch.rand(128, 5, generator=g).cuda()
    Y = T
    h = HOGP({'fidelity_shapes': [torch.Size(shape[1:])]}).double().cuda()

    def step():
        h.zero_grad(set_to_none=True)
        loss = h.compute_loss(x, Y)
        loss.backward()
        return loss

    sec = timeit(step)
    report('HOGP C4 compute_loss + backward (128x32x32x16)', sec, 4 * 16 * numel)
    import cProfile, pstats, io
    pr = cProfile.Profile()
    torch.cuda.synchronize()
    pr.enable()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    pr.disable()
    sio = io.StringIO()
    pstats.Stats(pr, stream=sio).sort_stats('cumulative').print_stats(28)
    print(sio.getvalue()[:6000])
    sec = timeit(lambda: h.compute_loss(x, Y))
    report('HOGP C4 compute_loss only', sec, 2 * 16 * numel)
    xs = torch.rand(32, 5, generator=g).cuda()
    sec = timeit(lambda: h.forward(xs))
    report('HOGP C4 forward (N*=32)', sec)

if not a.only or 'c3' in a.only:
    N = 128
    yl = torch.randn(N, 1024, generator=g).cuda()
    yh = torch.randn(N, 4096, generator=g).cuda()
    x = torch.rand(N, 5, generator=g).cuda()
    tlin = Tensor_linear([1024], [4096]).double().cuda()
    sec = timeit(lambda: tlin(yl))
    report('Tensor_linear 1024->4096, N=128 (weights 32 MiB)', sec, 8 * (N * 1024 + N * 4096 + 1024 * 4096), 2 * N * 1024 * 4096)
    m = cigp(ARDKernel(5), 1.0).cuda()

    def step3():
        m.zero_grad(set_to_none=True); tlin.zero_grad(set_to_none=True)
        res = yh - tlin(yl)
        loss = -m.negative_log_likelihood(x, res)
        loss.backward()
        return loss

    sec = timeit(step3)
    report('C3 top-fidelity step: residual + cigp NLL(D=4096, N=128) + backward', sec)
    for N3, D3 in ((512, 256), (256, 1024)):
        x3 = torch.rand(N3, 5, generator=g).cuda(); y3 = torch.randn(N3, D3, generator=g).cuda()

        def step3b():
            m.zero_grad(set_to_none=True)
            loss = -m.negative_log_likelihood(x3, y3)
            loss.backward()

        sec = timeit(step3b)
        report(f'cigp NLL+grad N={N3}, D={D3}', sec, None, N3 ** 3 + 4 * N3 * N3 * (5 + D3))

if a.json:
    json.dump({'hbm_peak_gbs': HBM, 'rows': rows}, open(a.json, 'w'), indent=1)
