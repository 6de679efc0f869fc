"""Device-side training loop.

The reference trains every model with the same Python loop (GaussianProcess/cigp_v10.py:160-175,
FidelityFusion_Models/CIGAR.py:96-133, AR_autoRegression.py:120-140, MFGP_ver2023May/mfgp_demo.py:40-48):

    optimizer = torch.optim.Adam(model.parameters(), lr=lr_init)
    for i in range(max_iter):
        optimizer.zero_grad(); loss = -model.negative_log_likelihood(x, y); loss.backward(); optimizer.step()

At the reference's real sizes (N = 100-300) an iteration is a handful of microsecond-scale kernels and the loop is
bound by Python, the autograd engine and launch latency.  Two pieces remove that:

* `FusedAdam`   - drop-in for `torch.optim.Adam` (same constructor keywords, `step()`, `zero_grad()`, per-parameter
                  step counts, parameters without a gradient are skipped).  The update of ALL parameter tensors is one
                  launch of `ffgp_adam_step_f64`; the step counters and the loss history stay on the device.
* `GraphedTrainer` - captures ONE iteration (zero the static gradients -> loss -> backward -> FusedAdam.step) into a
                  CUDA graph and replays it: no Python, no allocator and one `cudaGraphLaunch` per epoch.

Both run the same kernels as the eager path, so the parameter trajectory is the eager one bit for bit.
"""
from __future__ import annotations

import torch

from . import _lib as B
from . import ops


def _bump_versions(params):
    """The kernels write parameter memory behind autograd's back: advance the tensors' version counters, which is
    what an in-place torch op would have done (ops.state_token keys the resident factorisation on them, and autograd
    uses them to detect stale saved tensors)."""
    for p in params:
        torch.autograd.graph.increment_version(p)


class FusedAdam(torch.optim.Optimizer):
    """`torch.optim.Adam(params, lr, betas, eps, maximize=...)` with the update fused into one CUDA launch.

    weight_decay / amsgrad are not used by the reference and not supported.  Parameters must be contiguous fp64 CUDA
    tensors (the harness convention of SURVEY.md 8c: `.double()` models); anything else raises.
    `step(loss=...)` additionally records `loss` into a device-side history (`losses()`), replacing the
    `loss.item()` read of the reference's loops."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, maximize=False,
                 history=0):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError('FusedAdam: weight_decay / amsgrad are not supported (unused by the reference)')
        if not 0.0 <= lr or not 0.0 <= eps or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError('FusedAdam: bad hyper-parameter')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, maximize=maximize))
        self._tables = {}
        self._hist_cap = int(history)
        self._hist = None

    def _state_for(self, p):
        st = self.state[p]
        if not st:
            if not (p.is_cuda and p.dtype == torch.float64 and p.is_contiguous()):
                raise TypeError('FusedAdam needs contiguous fp64 CUDA parameters (model.double().cuda()); '
                                f'got {p.dtype} on {p.device}')
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st['step'] = torch.zeros(1, dtype=torch.float64, device=p.device)
        return st

    def _table(self, active):
        """Device pointer table of the tensors updated together; cached per (param, grad) address set so that a
        steady-state step allocates nothing and copies nothing (required under CUDA-graph capture)."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in active)
        hit = self._tables.get(key)
        if hit is None:
            for p in active:                           # validate before touching the CUDA runtime: no CPU fallback
                self._state_for(p)
                if not (p.grad.is_cuda and p.grad.dtype == torch.float64 and p.grad.is_contiguous()):
                    raise TypeError('FusedAdam needs contiguous fp64 CUDA gradients')
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError('FusedAdam: the set of (parameter, gradient) buffers changed during CUDA-graph capture; '
                                   'run one eager iteration first (GraphedTrainer does)')
            ptrs = []
            for p in active:
                st = self._state_for(p)
                ptrs += [p.data_ptr(), p.grad.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(),
                         st['step'].data_ptr()]
            dev = active[0].device
            hit = (torch.tensor(ptrs, dtype=torch.int64, device=dev),
                   torch.tensor([p.numel() for p in active], dtype=torch.int32, device=dev), list(active))
            if len(self._tables) > 16:                 # eager loops that re-allocate gradients every iteration
                self._tables.clear()
            self._tables[key] = hit
        return hit

    @torch.no_grad()
    def step(self, closure=None, loss=None):
        out = None
        if closure is not None:
            with torch.enable_grad():
                out = closure()
        L = B.lib()
        first = True
        for group in self.param_groups:
            active = [p for p in group['params'] if p.grad is not None]
            if not active:
                continue
            table, sizes, _ = self._table(active)
            loss_p = hist_p = None
            if first and loss is not None and self._hist_cap > 0:
                if self._hist is None:
                    self._hist = torch.zeros(self._hist_cap, dtype=torch.float64, device=active[0].device)
                lt = loss.detach().reshape(-1)
                if lt.dtype != torch.float64 or not lt.is_cuda:
                    raise TypeError('FusedAdam.step(loss=...): loss must be an fp64 CUDA tensor')
                loss_p, hist_p = B.ptr(lt), B.ptr(self._hist)
            b1, b2 = group['betas']
            rc = L.ffgp_adam_step_f64(B.ptr(table), B.ptr(sizes), len(active), float(group['lr']), float(b1), float(b2),
                                      float(group['eps']), int(bool(group['maximize'])), loss_p, hist_p, self._hist_cap,
                                      B.stream_ptr())
            B.check(rc, 'ffgp_adam_step_f64')
            _bump_versions(active)
            first = False
        return out

    def losses(self, n=None):
        """The recorded loss history (device tensor): entry i is the loss passed to the (i+1)-th step."""
        if self._hist is None:
            return torch.zeros(0, dtype=torch.float64)
        return self._hist if n is None else self._hist[:n]

    def reset_state(self):
        """Zero the moments, the step counters and the loss history in place (addresses are kept)."""
        for st in self.state.values():
            for k in ('exp_avg', 'exp_avg_sq', 'step'):
                if k in st:
                    st[k].zero_()
        if self._hist is not None:
            self._hist.zero_()


class GraphedTrainer:
    """One training iteration of the reference's loop, captured in a CUDA graph and replayed.

        trainer = GraphedTrainer(lambda: -model.negative_log_likelihood(x, y), model.parameters(), lr=1e-2)
        trainer.run(300)                 # 300 x cudaGraphLaunch, nothing else on the host
        curve = trainer.losses()         # [300] device tensor

    `loss_fn` must be a pure function of the parameters and of tensors whose storage does not change between
    iterations (the training data).  Construction runs `warmup` eager iterations on a side stream (they create the
    library's streams, workspace and pointer tables and discover which parameters receive gradients) and then restores
    parameters and optimiser state, so the trajectory starts from the values the caller passed in.
    Non-positive-definite failures are reported by `check()` / `run(check=True)` (the status word cannot be read while
    the graph is being captured or replayed without a host synchronisation)."""

    def __init__(self, loss_fn, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, warmup=2, history=4096):
        self.loss_fn = loss_fn
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('GraphedTrainer: no trainable parameters')
        if not all(p.is_cuda for p in self.params):
            raise B.FFGPError('GraphedTrainer needs CUDA parameters (no CPU fallback)')
        self.opt = FusedAdam(self.params, lr=lr, betas=betas, eps=eps, history=history)
        self.steps = 0
        snapshot = [p.detach().clone() for p in self.params]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            # iteration 0: gradients start as None -> afterwards exactly the parameters the loss reaches have one
            for p in self.params:
                p.grad = None
            self.loss_fn().backward()
            self.active = [p for p in self.params if p.grad is not None]
            for p in self.active:                       # static gradient buffers: backward accumulates in place
                p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
            for _ in range(max(1, warmup)):
                self._iteration()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        ops.flush_deferred_checks()
        with torch.no_grad():
            for p, s in zip(self.params, snapshot):
                p.copy_(s)
        self.opt.reset_state()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()
        # (capture does not execute: parameters and optimiser state are still the restored ones)
        # factorisation status words of the captured launches: they live in the graph's pool and are re-written by
        # every replay
        self._infos, ops._deferred_info[:] = list(ops._deferred_info), []

    def _iteration(self):
        for p in self.active:
            p.grad.zero_()
        loss = self.loss_fn()
        loss.backward()
        self.opt.step(loss=loss)

    def run(self, iterations, check=True):
        for _ in range(int(iterations)):
            self.graph.replay()
        self.steps += int(iterations)
        _bump_versions(self.active)                # a replay does not run the Python side of FusedAdam.step
        if check:
            self.check()
        return self

    def check(self):
        """Host synchronisation + the deferred positive-definiteness check of the last replay."""
        torch.cuda.synchronize()
        for info, what in self._infos:
            ops.check_info(info, what)

    def losses(self):
        return self.opt.losses(min(self.steps, self.opt._hist_cap))


def train(loss_fn, params, max_iter, lr=1e-3, graphed=True, **adam_kw):
    """The reference's training loop.  graphed=True: GraphedTrainer; False: the eager loop with FusedAdam.
    Returns the loss curve as a device tensor [max_iter]."""
    params = [p for p in params]
    if graphed:
        return GraphedTrainer(loss_fn, params, lr=lr, history=max_iter, **adam_kw).run(max_iter).losses()
    opt = FusedAdam(params, lr=lr, history=max_iter, **adam_kw)
    for _ in range(max_iter):
        opt.zero_grad(set_to_none=False)           # gradient buffers keep their addresses: the pointer table is reused
        loss = loss_fn()
        loss.backward()
        opt.step(loss=loss)
    return opt.losses(max_iter)
