"""S = M^T M (lower tiles, K range [i0, N)) for a batch of N = 512 problems: the largest launch of a batched evaluation.
   FFGP_PERSIST_LOWER=0|1 python tools/lauum_bench.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fidelityfusion_b200 import _lib as B
L = B.lib(); st = B.stream_ptr()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = 512
M = torch.tril(torch.randn(batch, n, n, dtype=torch.float64, device='cuda'))
S = torch.zeros(batch, n, n, dtype=torch.float64, device='cuda')
def f(): return L.ffgp_gemm_f64(0, 0, B.ptr(M), n, n * n, B.ptr(M), n, n * n, B.ptr(S), n, n * n, n, n, n, 1.0, 0.0, 1, 4, batch, st)
for _ in range(3): assert f() == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
ref = torch.tril(M[:2].transpose(1, 2) @ M[:2])
err = float((torch.tril(S[:2]) - ref).abs().max() / ref.abs().max())
print(f'PERSIST_LOWER={os.environ.get("FFGP_PERSIST_LOWER", "0")} batch {batch}: {ms:.3f} ms per launch = {ms / batch * 1e3:.3f} us per problem, '
      f'{batch * n ** 3 / 3 * 2 / 2 / ms / 1e9:.2f} TFLOP/s on N^3/3, max rel err {err:.1e}')
