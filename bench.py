#!/usr/bin/env python
"""Benchmark of the GP hot path (BASELINE.json metric: "GP NLL+grad evals/sec (fp64, N=8192, d=16) & batched GPs/sec;
% FP64 roofline").

  python bench.py --gpus N --steps K --warmup W          # our CUDA path (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --steps K --warmup W  # the reference's CPU path (oracle port) on the host cores

A step = ONE NLL+gradient evaluation of the single-fidelity ARD-RBF cigp of BASELINE config 2 (synthetic N=8192, d=16,
D=1, fp64).  The large-N factorisation is not distributed (north_star): with N>1 GPUs every rank evaluates its own
replica (e.g. a hyper-parameter restart) - "replicas only", weak scaling, no data-path collective.  The batched config
(4096 independent GPs of N=512, d=8: BASELINE config 5) IS sharded across ranks with one all-gather and is reported in
the `batched` object of the same JSON line.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_C2, D_C2 = 8192, 16
F_ALG_C2 = N_C2 ** 3 + 4 * N_C2 ** 2 * D_C2 + 4 * N_C2 ** 2 * 1            # 5.543e11 FLOP / eval
B_C5, N_C5, D_C5, NS_C5 = 4096, 512, 8, 64
F_ALG_C5 = N_C5 ** 3 + 4 * N_C5 ** 2 * D_C5 + 4 * N_C5 ** 2 * 1            # 1.437e8 FLOP / GP
# dram__bytes_read.sum + dram__bytes_write.sum of ONE S = M^T M launch of the dominant kernel at N=8192 (ncu,
# profiles/r01_metrics_c2_v1.txt: 2 launches read 2794 MB and wrote 460 MB); algorithmic bytes 8 N^2 = 537 MB
TRAFFIC_LAUUM_BYTES = 1.788e9     # dram read 1.561 GB + write 0.227 GB per launch, profiles/r02_ncu_gemm_tma_lauum_final.txt


def c2_inputs(torch):
    """SURVEY.md 8(d) C2 recipe (seed 0)."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N_C2, D_C2, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(1, keepdim=True)) + 0.1 * torch.randn(N_C2, 1, generator=g, dtype=torch.float64)
    return x, y


def rowdot(x, w):
    """x [B,n,d] . w [B,d,1] as d sequential element-wise multiply-adds: the same bits in every process on every host.
    A CPU matmul is not (MKL chooses its summation order by the buffers' alignment: 1 process in ~24 on the GPU box's host
    produced a `y` that differed in the last bits, and with it the checksum - tools/launches_c5.py prints the evidence)."""
    acc = x[..., 0:1] * w[:, 0:1, :]
    for k in range(1, x.shape[-1]):
        acc = acc + x[..., k:k + 1] * w[:, k:k + 1, :]
    return acc


def c5_inputs(torch, lo, hi):
    g = torch.Generator().manual_seed(5000)
    x = torch.rand(B_C5, N_C5, D_C5, generator=g, dtype=torch.float64)
    w = torch.randn(B_C5, D_C5, 1, generator=g, dtype=torch.float64)
    y = torch.sin(3 * rowdot(x, w)) + 0.05 * torch.randn(B_C5, N_C5, 1, generator=g, dtype=torch.float64)
    ls = torch.exp(torch.rand(B_C5, D_C5, generator=g, dtype=torch.float64) * 2 - 1)
    lb = torch.rand(B_C5, generator=g, dtype=torch.float64) * 3
    xs = torch.rand(B_C5, NS_C5, D_C5, generator=g, dtype=torch.float64)
    sl = slice(lo, hi)
    return x[sl], y[sl], ls[sl], torch.ones(hi - lo, dtype=torch.float64), lb[sl], xs[sl]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        # ONE long-lived nvidia-smi in loop mode (-lms 100): a fresh process per sample costs ~250 ms and a 0.5 s timed
        # region would see two samples
        try:
            self._proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i',
                                           str(self.index), '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                          text=True)
        except Exception:
            return
        for line in self._proc.stdout:
            if self._stop.is_set():
                break
            line = line.strip()
            if line and self._armed.is_set():
                self.samples.append([f.strip() for f in line.split(',')])

    def arm(self):
        """Start recording (call right before the timed region; the process is already warm)."""
        self._armed.set()

    def __enter__(self):
        self._armed = threading.Event()
        self._proc = None
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._proc is not None:
            try:
                self._proc.terminate()
            except Exception:
                pass
        self._t.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith('active') for s in self.samples if len(s) > 3 + i)]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(self.samples)}


def _reference_tree():
    """A staged copy of the UNMODIFIED reference (baseline/_ref: git-ignored scratch written by tools/stage_reference.sh,
    it travels to the GPU box).  /root/reference itself is never read here."""
    p = os.path.join(ROOT, 'baseline', '_ref')
    return p if os.path.isdir(os.path.join(p, 'GaussianProcess')) else None


_REF_MODULES = None


def _import_reference_cigp(torch):
    """cigp + ARDKernel of the staged reference tree (its own files, unmodified), or None."""
    global _REF_MODULES
    if _REF_MODULES is not None:
        return _REF_MODULES or None
    ref = _reference_tree()
    _REF_MODULES = False
    if ref is None:
        return None
    try:
        import contextlib, io, types
        for name in ('matplotlib', 'matplotlib.pyplot'):          # imported at module top for the __main__ demos only
            if name not in sys.modules:
                try:
                    __import__(name)
                except Exception:
                    sys.modules[name] = types.ModuleType(name)
        if 'matplotlib.pyplot' in sys.modules and not hasattr(sys.modules['matplotlib'], 'pyplot'):
            sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
        sys.path.insert(0, ref)
        with contextlib.redirect_stdout(io.StringIO()):
            from GaussianProcess.cigp_v10 import cigp
            from GaussianProcess.kernel import ARDKernel
        if not cigp.__module__.startswith('GaussianProcess') or 'fidelityfusion_b200' in sys.modules[cigp.__module__].__file__:
            return None
        _REF_MODULES = (cigp, ARDKernel)
    except Exception:
        _REF_MODULES = False
    return _REF_MODULES or None


def cpu_eval_c2(torch, threads):
    """One NLL+grad eval of C2 on the host cores exactly as the reference performs it (CIGAR.py:100-105:
    loss = -gpr.negative_log_likelihood(x, y); loss.backward()).  With a staged reference tree (baseline/_ref) this IS the
    reference's own cigp + ARDKernel ('kind': 'reference'); otherwise the oracle restatement of the same torch calls
    (cdist kernel, torch.linalg.cholesky, triangular solve, autograd backward; 'kind': 'port')."""
    torch.set_num_threads(threads)
    x, y = c2_inputs(torch)
    ref = _import_reference_cigp(torch)
    if ref is not None:
        cigp, ARDKernel = ref
        m = cigp(ARDKernel(D_C2), 1.0).double()
        t0 = time.perf_counter()
        loss = -m.negative_log_likelihood(x, y)
        loss.backward()
        dt = time.perf_counter() - t0
        cpu_eval_c2.last_grads = {'length_scales': m.kernel.length_scales.grad, 'signal_variance': m.kernel.signal_variance.grad,
                                  'log_beta': m.log_beta.grad}
        cpu_eval_c2.kind = 'reference'
        return dt, float(loss.item())
    from oracle import ff_oracle as O
    ls, sv, lb = torch.ones(D_C2, dtype=torch.float64), torch.ones(1, dtype=torch.float64), torch.ones(1, dtype=torch.float64)
    t0 = time.perf_counter()
    loss, grads = O.cigp_ard_nll_and_grads(x, y, ls, sv, lb)
    cpu_eval_c2.last_grads = grads
    cpu_eval_c2.kind = 'port'
    return time.perf_counter() - t0, loss


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port, all host threads)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    torch.set_default_dtype(torch.float64)
    threads = os.cpu_count() or 1
    budget = 240.0
    t_start = time.perf_counter()
    times = []
    for i in range(args.warmup + args.steps):
        dt, loss = cpu_eval_c2(torch, threads)
        if i >= args.warmup:
            times.append(dt)
        elif i == 0 and dt * (args.warmup + args.steps) > budget:
            # keep the whole run within a few minutes: drop the remaining warm-up evals (each eval is seconds long,
            # far above any cold-start effect)
            args.warmup = 1
        if time.perf_counter() - t_start > budget and len(times) >= 1:
            break
    ms = 1e3 * sum(times) / len(times)
    val = 1e3 / ms
    line = {
        'impl': 'reference', 'metric': 'GP NLL+grad evals/sec (fp64, N=8192, d=16)', 'value': val, 'unit': 'evals/s',
        'n_gpus': args.gpus, 'steps': len(times), 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C2: single-fidelity ARD-RBF cigp NLL+grad, N=8192, d=16, D=1, fp64 (SURVEY 8d recipe, seed 0)'},
        'cpu_baseline': {'value': val, 'unit': 'evals/s', 'cores': threads, 'kind': cpu_eval_c2.kind,
                         'sample': f'{len(times)} full NLL+grad evals of the C2 workload (N=8192) on the host, torch {torch.__version__} CPU/MKL; '
                                   + ('the unmodified reference cigp + ARDKernel from baseline/_ref' if cpu_eval_c2.kind == 'reference'
                                      else 'oracle restatement of the reference torch calls')},
        'e2e': {'value': val, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'nll': loss,
    }
    print(json.dumps(line), flush=True)


def kron_numbers(torch):
    """Kronecker / coupling half of the path (BASELINE configs 3 and 4): HBM-bound mode products and the whole HOGP
    loss+gradient step at the C4 shape 128 x 32 x 32 x 16.  CUDA events, median of 7, L2 evicted by a 512 MB read
    between repetitions.  Algorithmic bytes per mode product: 16 B per tensor element + the factor."""
    from fidelityfusion_b200 import tensorly_compat as tl
    from fidelityfusion_b200.MFGP_ver2023May import HOGP
    hbm = 6457.7
    try:
        hbm = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    flush = torch.zeros(512 << 20, dtype=torch.uint8, device='cuda').view(torch.int64)

    def med(fn, reps=7):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2] * 1e-3

    g = torch.Generator().manual_seed(4)
    out = {'hbm_peak_gbs': hbm, 'peak_source': 'MEASURED_PEAKS.json hbm_gbs (burst copy figure)'}
    for tag, n in (('c4_128x32x32x16', 128), ('large_1024x32x32x16', 1024)):
        T = torch.randn(n, 32, 32, 16, generator=g, dtype=torch.float64).cuda()
        row = {}
        for mode in (1, 3):
            I = T.shape[mode]
            U = torch.randn(I, I, generator=g, dtype=torch.float64).cuda()
            sec = med(lambda: tl._mode_dot_raw(T, U, mode, False))
            gbs = (16 * T.numel() + 8 * I * I) / sec * 1e-9
            row[f'mode_dot_mode{mode}'] = {'us': sec * 1e6, 'GBps': gbs, 'hbm_frac': gbs / hbm}
        sec = med(lambda: tl._mode_gram_raw(T, T, 1))
        row['mode_gram_mode1'] = {'us': sec * 1e6, 'GBps': 16 * T.numel() / sec * 1e-9, 'hbm_frac': 16 * T.numel() / sec * 1e-9 / hbm}
        sec = med(lambda: T.clone())
        row['torch_clone_same_bytes'] = {'us': sec * 1e6, 'GBps': 16 * T.numel() / sec * 1e-9}
        out[tag] = row
        del T
    x = torch.rand(128, 5, generator=g, dtype=torch.float64).cuda()
    Y = torch.randn(128, 32, 32, 16, generator=g, dtype=torch.float64).cuda()
    h = HOGP({'fidelity_shapes': [torch.Size([32, 32, 16])]}).double().cuda()

    def step():
        h.zero_grad(set_to_none=True)
        h.compute_loss(x, Y).backward()

    out['hogp_c4_loss_grad_ms'] = med(step) * 1e3
    try:      # the same step as ONE CUDA-graph replay (+ fused Adam): what the device-side training loop costs per epoch
        from fidelityfusion_b200.training import GraphedTrainer
        h2 = HOGP({'fidelity_shapes': [torch.Size([32, 32, 16])]}).double().cuda()
        tr = GraphedTrainer(lambda: h2.compute_loss(x, Y), h2.parameters(), lr=1e-3, history=64)
        tr.run(3)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tr.run(20, check=False)
        torch.cuda.synchronize()
        out['hogp_c4_graph_replay_ms_per_epoch'] = (time.perf_counter() - t0) / 20 * 1e3
        tr.check()
    except Exception as e:
        out['hogp_c4_graph_replay_ms_per_epoch'] = repr(e)[:200]
    K = torch.exp(-0.5 * torch.cdist(x, x) ** 2)
    out['eigh_n128_ms'] = med(lambda: tl.eigh(K)) * 1e3
    out['eigh_n128_cusolver_ms'] = med(lambda: torch.linalg.eigh(K)) * 1e3        # the library kernel it replaces, same box
    for n_small in (32, 16):
        Ks = K[:n_small, :n_small].contiguous()
        out[f'eigh_n{n_small}_ms'] = med(lambda: tl.eigh(Ks)) * 1e3
        out[f'eigh_n{n_small}_cusolver_ms'] = med(lambda: torch.linalg.eigh(Ks)) * 1e3
    return out


def training_numbers(torch, iters=100):
    """Device-side training loop at the reference's real size (N=300, d=2: FidelityFusion_Models/log/AR/
    train.log, 11 ms/epoch on the reference's CPU): ms per epoch (zero_grad -> NLL -> backward -> Adam.step) of the
    eager loop with torch.optim.Adam and of the CUDA-graph replay with the fused Adam kernel."""
    from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    from fidelityfusion_b200.training import GraphedTrainer
    g = torch.Generator().manual_seed(300)
    x = torch.rand(300, 2, generator=g, dtype=torch.float64).cuda()
    y = (torch.sin(3 * x.sum(1, keepdim=True)) + 0.1 * torch.randn(300, 1, generator=g, dtype=torch.float64).cuda())
    m = cigp(ARDKernel(2), 1.0).double().cuda()
    opt = torch.optim.Adam(m.parameters(), lr=0.01)

    def it():
        opt.zero_grad(); loss = -m.negative_log_likelihood(x, y); loss.backward(); opt.step()
    for _ in range(5):
        it()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(iters):
        it()
    torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / iters * 1e3
    m2 = cigp(ARDKernel(2), 1.0).double().cuda()
    tr = GraphedTrainer(lambda: -m2.negative_log_likelihood(x, y), m2.parameters(), lr=0.01, history=2 * iters + 8)
    tr.run(5); torch.cuda.synchronize(); t0 = time.perf_counter()
    tr.run(iters, check=False); torch.cuda.synchronize(); graphed = (time.perf_counter() - t0) / iters * 1e3
    tr.check()
    return {'config': 'cigp + ARDKernel, N=300, d=2, D=1, Adam', 'eager_torch_adam_ms_per_epoch': eager,
            'graph_replay_fused_adam_ms_per_epoch': graphed, 'final_loss': float(tr.losses()[-1])}


def reference_cuda_numbers(torch):
    """INFORMATIONAL ("these library calls are the existing kernels the new sm_100a kernels must beat on the
    same box"): what the reference itself does when a user moves its tensors to the GPU - the oracle port executed on
    CUDA tensors, i.e. torch.cdist / cuBLAS, cuSOLVER potrf + trsm, syevd and autograd through them.  No parity or
    roofline claim rides on these numbers; they are the library baseline beside ours, same box, same run."""
    from oracle import ff_oracle as O
    out = {'what': 'oracle port (the reference torch call sequence) on CUDA tensors: cuBLAS / cuSOLVER / autograd'}
    torch.set_default_device('cuda')          # the reference builds torch.eye() / constants on the default device
    try:
        return _reference_cuda_body(torch, O, out)
    finally:
        torch.set_default_device('cpu')


def _reference_cuda_body(torch, O, out):

    def timed(fn, reps):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # C2: one NLL + gradient evaluation, N = 8192, d = 16
    with torch.device('cpu'):
        xc_, yc_ = c2_inputs(torch)
    x, y = xc_.cuda(), yc_.cuda()
    ls, sv, lb = (torch.ones(D_C2, device='cuda'), torch.ones(1, device='cuda'), torch.ones(1, device='cuda'))
    ms = timed(lambda: O.cigp_ard_nll_and_grads(x, y, ls, sv, lb), 3)
    out['c2_ms_per_eval'] = ms
    out['c2_evals_per_s'] = 1e3 / ms
    del x, y
    torch.cuda.empty_cache()
    # C5: 64 of the 4096 problems, sequentially like the reference's Python loops (v1/CFKG.py:124-129)
    with torch.device('cpu'):
        c5 = c5_inputs(torch, 0, 64)
    bx, by, bls, bsv, blb, bxs = [t.cuda() for t in c5]

    def c5_loop():
        for b in range(64):
            O.cigp_ard_nll_and_grads(bx[b], by[b], bls[b], bsv[b:b + 1], blb[b:b + 1])
            O.cigp_ard_predict(bx[b], by[b], bxs[b], bls[b], bsv[b:b + 1], blb[b:b + 1])
    ms = timed(c5_loop, 2)
    out['c5_gps_per_s_sequential_loop'] = 64 / ms * 1e3
    # the same 64 problems batched through torch's own batched kernels (what a user could write with torch alone)
    def c5_batched():
        l = bls.clone().requires_grad_(True); s_ = bsv.clone().requires_grad_(True); b_ = blb.clone().requires_grad_(True)
        ell = l.abs() + 1e-9
        K = s_.abs().view(-1, 1, 1) * torch.exp(-0.5 * torch.cdist(bx / ell.unsqueeze(1), bx / ell.unsqueeze(1)) ** 2)
        Sig = K + (torch.exp(-b_) + 1e-6).view(-1, 1, 1) * torch.eye(N_C5, device='cuda')
        Lc = torch.linalg.cholesky(Sig)
        gam = torch.linalg.solve_triangular(Lc, by, upper=False)
        nll = 0.5 * (gam ** 2).sum((1, 2)) + torch.log(torch.diagonal(Lc, dim1=1, dim2=2)).sum(1)
        nll.sum().backward()
    ms = timed(c5_batched, 3)
    out['c5_gps_per_s_torch_batched_nll_grad_only'] = 64 / ms * 1e3
    del bx, by
    # C4: one HOGP loss + gradient, 128 x 32 x 32 x 16 (eigh through cuSOLVER syevd, autograd through it)
    with torch.device('cpu'):
        g = torch.Generator().manual_seed(4)
        xk = torch.rand(128, 5, generator=g, dtype=torch.float64)
        Y = torch.randn(128, 32, 32, 16, generator=g, dtype=torch.float64)
    xk, Y = xk.cuda(), Y.cuda()
    grids = [xk] + [torch.arange(s_, dtype=torch.float64, device='cuda').reshape(-1, 1) for s_ in (32, 32, 16)]

    def hogp():
        p = [torch.ones(2, device='cuda', requires_grad=True) for _ in range(4)]
        nz = torch.ones(1, device='cuda', requires_grad=True)
        Ks = [O.se_kernel(grids[k], grids[k], p[k][0], p[k][1], False) for k in range(4)]
        loss, _, _ = O.hogp_loss(Ks, 1.0 / nz, Y)
        loss.backward()
    out['c4_hogp_loss_grad_ms'] = timed(hogp, 5)
    return out


def measure_dgemm_peak(torch):
    """FP64 roofline denominator: cuBLAS DGEMM 8192^3 through torch.matmul, best of 10 (MEASURED_PEAKS.json
    has no fp64 entry)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2 * n ** 3 / best * 1e-9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--skip-batched', action='store_true')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--skip-kron', action='store_true', help='skip the secondary kron / training / reference_cuda blocks')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else max(args.warmup, 0)
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device - the product path has no CPU fallback')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    torch.set_default_dtype(torch.float64)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    from fidelityfusion_b200 import _lib
    from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
    from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
    from fidelityfusion_b200.batched import batched_cigp_eval, check_batch_info, shard_range, sharded_cigp_eval
    lib = _lib.lib()                                    # fails loudly if the extension is missing

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak_tflops = measure_dgemm_peak(torch) if rank == 0 else 0.0

    # ------------------------------------------------------------------ C2: one replica per rank
    xh, yh = c2_inputs(torch)
    xh, yh = xh.pin_memory(), yh.pin_memory()
    model = cigp(ARDKernel(D_C2), 1.0).cuda()
    model.factor_cache = None
    xd, yd = xh.cuda(), yh.cuda()

    def one_eval(x, y):
        model.zero_grad(set_to_none=True)
        loss = -model.negative_log_likelihood(x, y)      # the reference's training-step idiom, CIGAR.py:100-105
        loss.backward()
        return loss

    clk = ClockSampler(local)
    clk.__enter__()                                      # nvidia-smi loop process starts now, recording is armed below
    for _ in range(args.warmup):
        one_eval(xd, yd)
    barrier()
    launches0 = lib.ffgp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk.arm()
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = one_eval(xd, yd)
    e1.record()
    barrier()
    launches = int(lib.ffgp_launch_count() - launches0)
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    nll_val = float(loss.item())
    value = world * 1e3 / ms_step
    gpu_grads = {'length_scales': model.kernel.length_scales.grad.detach().cpu().clone(),
                 'signal_variance': model.kernel.signal_variance.grad.detach().cpu().clone(),
                 'log_beta': model.log_beta.grad.detach().cpu().clone()}

    # ------------------------------------------------------------------ e2e: host buffers in, scalars + gradients out
    def one_eval_e2e():
        x = xh.cuda(non_blocking=True)
        y = yh.cuda(non_blocking=True)
        l = one_eval(x, y)
        out = torch.cat([l.reshape(1), model.kernel.length_scales.grad, model.kernel.signal_variance.grad, model.log_beta.grad])
        return out.cpu()                                 # device -> host read of the step's result (sync)

    for _ in range(2):
        one_eval_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = one_eval_e2e()
    barrier()
    ms_e2e = max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps)
    e2e = {'value': world * 1e3 / ms_e2e, 'unit': 'evals/s', 'h2d_bytes_per_step': int(xh.numel() * 8 + yh.numel() * 8),
           'd2h_bytes_per_step': int(res.numel() * 8), 'ms_per_step': ms_e2e}
    clk.__exit__()                                       # clocks sampled over the device-timed and the e2e loops

    # ------------------------------------------------------------------ dominant kernel alone: the TMA + DMMA GEMM in its
    # S = M^T M launch shape (lower tiles, K range [i0, N)): N^3/3 algorithmic FLOP per launch, CUDA events
    kern = None
    if rank == 0:
        n = N_C2
        Mt = torch.tril(torch.randn(n, n, device='cuda'))
        St = torch.empty(n, n, device='cuda')
        def lauum():
            rc = lib.ffgp_gemm_f64(0, 0, _lib.ptr(Mt), n, 0, _lib.ptr(Mt), n, 0, _lib.ptr(St), n, 0, n, n, n, 1.0, 0.0, 1, 4, 1,
                                   _lib.stream_ptr())
            assert rc == 0
        for _ in range(3):
            lauum()
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(10):
            lauum()
        k1.record(); torch.cuda.synchronize()
        kms = k0.elapsed_time(k1) / 10
        kern = {'name': 'gemm_tma_kernel<0,0> (S = M^T M launch of the gradient, 2080 lower 128x128 tiles)',
                'flop_per_launch': n ** 3 / 3, 'ms_per_launch': kms, 'achieved': n ** 3 / 3 / kms * 1e-9}
        del Mt, St

    # ------------------------------------------------------------------ batched C5, sharded across ranks
    batched = None
    if not args.skip_batched:
        lo, hi = shard_range(B_C5, rank, world)
        bx, by, bls, bsv, blb, bxs = [t.cuda() for t in c5_inputs(torch, 0, B_C5)]

        def sweep():
            # rank-local block [lo,hi) + ONE packed all-gather.  check=False: the positive-definiteness status travels
            # with the results (no host sync per sweep) and is checked once after the timed region, below
            return sharded_cigp_eval(bx, by, bls, bsv, blb, bxs, check=False)

        for _ in range(3):
            sweep()
        barrier()
        nsw = 5
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(nsw):
            out = sweep()
        b1.record()
        barrier()
        ms_sweep = max_over_ranks(b0.elapsed_time(b1) / nsw)
        check_batch_info(out['info'])
        gps = B_C5 / ms_sweep * 1e3
        batched = {'metric': 'batched GPs/sec (NLL+grad+predict, 4096 x N=512, d=8, N*=64)', 'value': gps, 'unit': 'GPs/s',
                   'ms_per_sweep': ms_sweep, 'scaling': 'strong', 'per_rank_problems': hi - lo,
                   'collective': 'none' if world == 1 else 'one all_gather_into_tensor of [B/R, 1+10+64+64] f64 per sweep',
                   'tflops_alg': gps * F_ALG_C5 * 1e-12, 'nll_checksum': float(out['nll'].sum().item()),
                   'pd_check': 'status word gathered with the results, checked once after the timed sweeps (check=False API)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        threads = os.cpu_count() or 1
        dt, cpu_loss = cpu_eval_c2(torch, threads)
        cpu = {'value': 1.0 / dt, 'unit': 'evals/s', 'cores': threads, 'kind': cpu_eval_c2.kind,
               'sample': '1 full NLL+grad eval of the C2 workload (N=8192, d=16) with '
                         + ('the unmodified reference cigp + ARDKernel (baseline/_ref) ' if cpu_eval_c2.kind == 'reference'
                            else 'the oracle port of the reference ') + f'(torch {torch.__version__} CPU); nll={cpu_loss:.10f}',
               'gpu_vs_cpu_nll_rel_diff': abs(cpu_loss - nll_val) / abs(cpu_loss),
               # all 18 hyper-parameter gradients of the same eval, max |gpu - cpu| / max |cpu| per parameter tensor
               'gpu_vs_cpu_grad_rel_diff': max(
                   float((gpu_grads[k] - v).abs().max() / v.abs().max()) for k, v in cpu_eval_c2.last_grads.items())}

    achieved = F_ALG_C2 * (world * 1e3 / ms_step) * 1e-12 / world      # per-GPU TFLOP/s on algorithmic FLOPs
    line = {
        'metric': 'GP NLL+grad evals/sec (fp64, N=8192, d=16)', 'value': value, 'unit': 'evals/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C2: single-fidelity ARD-RBF cigp NLL+grad, N=8192, d=16, D=1, fp64 (SURVEY 8d recipe, seed 0)',
                   'parallelism': 'replicas only (one independent eval per GPU; a single large-N Cholesky is not distributed)',
                   'l2': 'working set 1.6 GB per eval >> 126 MB L2 (no flush needed)', 'nll': nll_val},
        'e2e': e2e,
        'gpu_launches': launches,
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tflops, 'unit': 'TFLOP/s',
                     'frac': achieved / peak_tflops if peak_tflops else None, 'traffic': TRAFFIC_LAUUM_BYTES,
                     'kernel': kern,
                     'note': 'FP64 tensor pipe (DMMA). achieved = F_alg (N^3 + 4N^2 d + 4N^2 D = 5.543e11 FLOP/eval, SURVEY 8d) / '
                             'CUDA-event time of the whole eval (all launches; the DMMA GEMM kernel is >90% of it). '
                             "'kernel' = the dominant kernel timed alone (its own achieved TFLOP/s on N^3/3 FLOP); 'traffic' = its DRAM bytes per launch from ncu. "
                             'peak = torch.matmul fp64 8192^3 best of 10 measured live in this run (MEASURED_PEAKS.json has '
                             'no fp64 entry); DMMA issue-bound peak measured at 37.1 TFLOP/s (profiles/r01_fp64_peak_microbench.txt)'},
        'clocks': clk.summary(),
    }
    if cpu is not None:
        line['cpu_baseline'] = cpu
    if world == 1 and not args.skip_kron:
        try:
            line['kron'] = kron_numbers(torch)
        except Exception as e:                          # secondary numbers must never cost the headline line
            line['kron'] = {'error': repr(e)[:300]}
        try:
            line['training'] = training_numbers(torch)
        except Exception as e:
            line['training'] = {'error': repr(e)[:300]}
        try:
            torch.cuda.empty_cache()
            line['reference_cuda'] = reference_cuda_numbers(torch)
        except Exception as e:
            line['reference_cuda'] = {'error': repr(e)[:300]}
    if batched is not None:
        batched['roofline_frac'] = batched['tflops_alg'] / world / peak_tflops if peak_tflops else None
        line['batched'] = batched
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
