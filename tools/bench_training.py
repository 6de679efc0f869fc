"""Epoch time of the reference's training loop (zero_grad -> NLL -> backward -> Adam.step) at the reference's real sizes:
CPU oracle + torch Adam | our ops + torch.optim.Adam (eager) | our ops + FusedAdam (eager) | GraphedTrainer (CUDA graph).
   python tools/bench_training.py [--iters 200]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
from fidelityfusion_b200.training import FusedAdam, GraphedTrainer
from oracle import ff_oracle as O

ap = argparse.ArgumentParser(); ap.add_argument('--iters', type=int, default=200)
a = ap.parse_args()
out = {}
for n, d, D in ((100, 2, 1), (300, 2, 1), (300, 5, 64), (1024, 8, 1)):
    g = torch.Generator().manual_seed(n)
    x = torch.rand(n, d, generator=g); y = torch.sin(3 * x.sum(1, keepdim=True)).repeat(1, D) + 0.1 * torch.randn(n, D, generator=g)
    xc, yc = x.cuda(), y.cuda()
    row = {}
    # CPU: the oracle restatement of the reference + torch Adam
    ls = torch.ones(d, requires_grad=True); sv = torch.ones(1, requires_grad=True); lb = torch.ones(1, requires_grad=True)
    opt = torch.optim.Adam([ls, sv, lb], lr=0.01)
    k = max(10, min(a.iters, 4000 // n))
    def cpu_it():
        opt.zero_grad(); loss = -O.cigp_log_likelihood(O.ard_kernel(x, x, ls, sv), lb, y); loss.backward(); opt.step()
    for _ in range(3): cpu_it()
    t0 = time.perf_counter()
    for _ in range(k): cpu_it()
    row['cpu_oracle_ms'] = (time.perf_counter() - t0) / k * 1e3
    def gpu_loop(make_opt, fused):
        m = cigp(ARDKernel(d), 1.0).cuda()
        o = make_opt(m.parameters())
        def it():
            o.zero_grad(set_to_none=False) if fused else o.zero_grad()
            loss = -m.negative_log_likelihood(xc, yc); loss.backward()
            o.step(loss=loss) if fused else o.step()
        for _ in range(5): it()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(a.iters): it()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / a.iters * 1e3
    row['eager_torch_adam_ms'] = gpu_loop(lambda p: torch.optim.Adam(p, lr=0.01), False)
    row['eager_fused_adam_ms'] = gpu_loop(lambda p: FusedAdam(p, lr=0.01, history=a.iters + 8), True)
    m = cigp(ARDKernel(d), 1.0).cuda()
    tr = GraphedTrainer(lambda: -m.negative_log_likelihood(xc, yc), m.parameters(), lr=0.01, history=2 * a.iters + 8)
    tr.run(5); torch.cuda.synchronize(); t0 = time.perf_counter()
    tr.run(a.iters, check=False); torch.cuda.synchronize()
    row['graphed_ms'] = (time.perf_counter() - t0) / a.iters * 1e3
    tr.check()
    out[f'N={n},d={d},D={D}'] = {k_: round(v, 4) for k_, v in row.items()}
    print(f'N={n} d={d} D={D}:', out[f'N={n},d={d},D={D}'], flush=True)
print(json.dumps({'training_epoch_ms': out, 'iters': a.iters, 'cpu_threads': torch.get_num_threads()}))
