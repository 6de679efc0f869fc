"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck): a batched C5-shaped evaluation
(NLL + gradient + prediction, 12 x N = 300 -> padded 384, exercising the base kernel, the TMA and cp.async GEMMs, the
vector kernels and the gradient contraction), a second batch of 8 x N = 256 through the fused 256-block kernel and the lower-super-tile kernel-matrix kernel, a single dense evaluation with D > 8 (GEMM right-hand sides), the
Kronecker path (mode products, mode Gram, core, one-sided Jacobi in 1-, 2- and 8-CTA clusters) and the row matcher."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from fidelityfusion_b200.batched import batched_cigp_eval
from fidelityfusion_b200.GaussianProcess.cigp_v10 import cigp
from fidelityfusion_b200.GaussianProcess.kernel import ARDKernel
from fidelityfusion_b200.MFGP_ver2023May import HOGP
from fidelityfusion_b200 import tensorly_compat as tl, data_match
g = torch.Generator().manual_seed(1)
B, n, d, ns = 12, 300, 8, 40
x = torch.rand(B, n, d, generator=g).cuda(); y = torch.randn(B, n, 1, generator=g).cuda()
out = batched_cigp_eval(x, y, (torch.rand(B, d, generator=g) + 0.5).cuda(), torch.ones(B).cuda(), torch.rand(B, generator=g).cuda(),
                        torch.rand(B, ns, d, generator=g).cuda())
print('batched nll', float(out['nll'].sum()))
# round 2 kernels: 8 problems of n = 256 (np = 256: one factor256_kernel launch per problem, kernel_matrix_sym128_kernel, the
# symmetric / triangular bodies of the rolled TMA GEMM main loop in S = M^T M) - NLL + gradient + prediction
B2, n2 = 8, 256
x2 = torch.rand(B2, n2, d, generator=g).cuda(); y2 = torch.randn(B2, n2, 1, generator=g).cuda()
out2 = batched_cigp_eval(x2, y2, (torch.rand(B2, d, generator=g) + 0.5).cuda(), torch.ones(B2).cuda(), torch.rand(B2, generator=g).cuda(),
                         torch.rand(B2, ns, d, generator=g).cuda(), acq=dict(kind='EI', f_best=0.2, xi=0.01))
print('batched f256 nll', float(out2['nll'].sum()), 'score', float(out2['score'].sum()))
# third session of round 2: acquisition kinds of the stand-alone kernel (incl. the single-fidelity UCB / PI)
from fidelityfusion_b200.MF_BayesianOptimization.Discrete.DMF_acq import acquisition
for kind in ('UCB', 'EI', 'PI', 'UCB_STD', 'PI_CDF'):
    mu_ = out2['mean'][..., 0].clone().requires_grad_(True); v_ = out2['var'].clone().requires_grad_(True)
    s_ = acquisition(mu_, v_, kind, f_best=0.2, beta=1.5, xi=0.01)
    if kind != 'PI_CDF':
        s_.sum().backward()
print('batched acq kinds ok')
m = cigp(ARDKernel(5), 1.0).cuda()
xx = torch.rand(200, 5, generator=g).cuda(); yy = torch.randn(200, 24, generator=g).cuda()
(-m.negative_log_likelihood(xx, yy)).backward()
mean, cov = m(xx, yy, torch.rand(9, 5, generator=g).cuda())
print('dense D=24 ok', float(mean.sum()))
h = HOGP({'fidelity_shapes': [torch.Size([32, 12])]}).double().cuda()
xk = torch.rand(128, 3, generator=g).cuda(); Y = torch.randn(128, 32, 12, generator=g).cuda()
h.compute_loss(xk, Y).backward()
u, v = h.forward(torch.rand(5, 3, generator=g).cuda())
print('hogp ok', float(u.sum()))
for nn in (7, 16, 31, 50, 128):
    K = torch.exp(-0.5 * torch.cdist(xk[:nn], xk[:nn]) ** 2)
    w, V = tl.eigh(K)
print('eigh ok', float(w.sum()))
a = torch.rand(500, 5, generator=g).cuda()
print('match', int((data_match.row_match(a, a[100:300]) >= 0).sum()))
torch.cuda.synchronize()
