"""n-mode products with tensorly's call signatures (`tensorly.tenalg.mode_dot / multi_mode_dot`,
`tucker_to_tensor`, `tensor_to_vec`), running on libffgp's DMMA mode-product kernels with autograd
(reference call sites: hogp.py:132,173-187,217,236; multiscale_coupling/matrix.py:73,81;
gp_computation_pack.py:157).  tensorly itself is an un-vendored dependency of the reference; the
definition implemented is the standard one, fold(M @ unfold(T, mode))."""
import ctypes
import math

import torch

from . import _lib as B
from . import ops


def _split_shape(shape, mode):
    outer = math.prod(shape[:mode])
    inner = math.prod(shape[mode + 1:])
    return outer, shape[mode], inner


def _mode_dot_raw(t, mat, mode, transpose):
    """t: contiguous fp64 CUDA tensor; mat: [J][I] (transpose=False) or [I][J] (transpose=True)."""
    L = B.lib()
    outer, I, inner = _split_shape(t.shape, mode)
    J = mat.shape[1] if transpose else mat.shape[0]
    assert (mat.shape[0] if transpose else mat.shape[1]) == I, (tuple(mat.shape), tuple(t.shape), mode)
    out = torch.empty(t.shape[:mode] + (J,) + t.shape[mode + 1:], dtype=torch.float64, device=t.device)
    rc = L.ffgp_mode_dot_f64(B.ptr(t), B.ptr(mat), B.ptr(out), outer, I, inner, J, int(transpose), B.stream_ptr())
    B.check(rc, 'ffgp_mode_dot_f64')
    return out


def _mode_gram_raw(X, Y, mode):
    """G[a][b] = sum_rest X[.., a, ..] * Y[.., b, ..] along `mode` (X, Y contiguous, same shape but for `mode`)."""
    L = B.lib()
    outer, Ja, inner = _split_shape(X.shape, mode)
    Jb = Y.shape[mode]
    G = torch.empty(Ja, Jb, dtype=torch.float64, device=X.device)
    sb = L.ffgp_mode_gram_scratch_bytes(outer, inner, Ja, Jb)
    scratch = torch.empty(sb, dtype=torch.uint8, device=X.device)
    rc = L.ffgp_mode_gram_f64(B.ptr(X), B.ptr(Y), B.ptr(G), outer, inner, Ja, Jb, B.ptr(scratch), sb, B.stream_ptr())
    B.check(rc, 'ffgp_mode_gram_f64')
    return G


class _ModeDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, mat, mode, transpose):
        tc, mc = ops._f64c(t), ops._f64c(mat)
        ctx.save_for_backward(tc, mc)
        ctx.meta = (mode, transpose, t.dtype, mat.dtype)
        return _mode_dot_raw(tc, mc, mode, transpose).to(t.dtype)

    @staticmethod
    def backward(ctx, go):
        tc, mc = ctx.saved_tensors
        mode, transpose, tdt, mdt = ctx.meta
        goc = ops._f64c(go)
        gt = gm = None
        if ctx.needs_input_grad[0]:
            gt = _mode_dot_raw(goc, mc, mode, not transpose).to(tdt)
        if ctx.needs_input_grad[1]:
            gm = _mode_gram_raw(goc, tc, mode)           # [J][I]
            if transpose:
                gm = gm.t()
            gm = gm.to(mdt)
        return gt, gm, None, None


def mode_dot(tensor, matrix_or_vector, mode, transpose=False):
    """tensorly.tenalg.mode_dot: matrix [J, I_mode] maps mode size I -> J; a vector [I_mode] contracts the mode."""
    m = matrix_or_vector
    if m.dim() == 1:
        out = _ModeDot.apply(tensor, m.unsqueeze(0), mode, False)
        return out.squeeze(mode)
    return _ModeDot.apply(tensor, m, mode, bool(transpose))


def multi_mode_dot(tensor, matrix_or_vec_list, modes=None, skip=None, transpose=False):
    """tensorly.tenalg.multi_mode_dot: successive mode products (modes default to 0..len-1)."""
    if modes is None:
        modes = list(range(len(matrix_or_vec_list)))
    res = tensor
    decrement = 0
    for i, (m, mode) in enumerate(zip(matrix_or_vec_list, modes)):
        if skip is not None and i == skip:
            continue
        res = mode_dot(res, m, mode - decrement, transpose=transpose)
        if m.dim() == 1:
            decrement += 1
    return res


def tucker_to_tensor(tucker_tensor):
    core, factors = tucker_tensor
    return multi_mode_dot(core, factors)


def tensor_to_vec(t):
    return t.reshape(-1)


def ones(shape, **kw):
    return torch.ones(shape, **{k: v for k, v in kw.items() if k in ('device', 'dtype')})


_backend = 'pytorch'


def set_backend(name, *args, **kwargs):
    """tensorly.set_backend: six reference files call `tensorly.set_backend('pytorch')` at import
    (gp_computation_pack.py:12, hogp.py:7, matrix.py:6, both hogp_simple.py:13, GAR_GeneralizedAutoAR.py:11).
    torch tensors are the only kind handled here, so anything else is an error rather than a silent no-op."""
    if name != 'pytorch':
        raise ValueError(f"fidelityfusion_b200.tensorly_compat computes on torch CUDA tensors only (backend 'pytorch'), got {name!r}")


def get_backend():
    return _backend


def tensor(data, **kw):
    return torch.as_tensor(data, **{k: v for k, v in kw.items() if k in ('device', 'dtype')})


def to_numpy(t):
    return t.detach().cpu().numpy()


def unfold(t, mode):
    """tensorly.unfold: mode-`mode` matricisation [I_mode, prod(rest)] (row-major over the remaining modes)."""
    return torch.movedim(t, mode, 0).reshape(t.shape[mode], -1)


def fold(unfolded, mode, shape):
    shape = list(shape)
    full = [shape[mode]] + shape[:mode] + shape[mode + 1:]
    return torch.movedim(unfolded.reshape(full), 0, mode)


# ---------------------------------------------------------------------------------------------
# Kronecker GP core: per-mode eigh -> T1 -> (A, core, sums) -> g, with the analytic gradient
# ---------------------------------------------------------------------------------------------
_side_streams = {}


def _eigh_launch(K, stream=None):
    """Enqueue the Jacobi eigensolver; returns (w, V, info) without synchronising.  `stream`: run the solve on that
    side stream (forked from / joined to the current one by the caller); every buffer is allocated on the current
    stream so its lifetime is tied to the consumer."""
    L = B.lib()
    Kc = ops._f64c(K).unsqueeze(0)
    n = Kc.shape[-1]
    dev = K.device
    w = torch.empty(1, n, dtype=torch.float64, device=dev)
    V = torch.empty(1, n, n, dtype=torch.float64, device=dev)
    info = torch.empty(1, dtype=torch.int32, device=dev)
    wsb = L.ffgp_syevj_workspace_bytes(n, 1)
    # concurrent solves need their own scratch; the shared cache serves the solve on the current stream
    ws = ops._ws_cache.get(wsb, dev) if stream is None else torch.empty(wsb, dtype=torch.uint8, device=dev)
    sp = B.stream_ptr() if stream is None else ctypes.c_void_p(stream.cuda_stream)
    rc = L.ffgp_syevj_f64(B.ptr(Kc), n, 1, B.ptr(w), B.ptr(V), B.ptr(ws), wsb, B.ptr(info), sp)
    B.check(rc, 'ffgp_syevj_f64')
    return w[0], V[0], info, (Kc, ws)


def _eigh_launch_all(Ks):
    """All per-mode eigensolves of one Kronecker objective.  They are latency-bound single-cluster kernels (n = 128:
    5 ms, n = 32: 0.4 ms), so the largest runs on the current stream and the others concurrently on side streams."""
    if len(Ks) == 1 or not Ks[0].is_cuda:
        return [_eigh_launch(K) for K in Ks]
    cur = torch.cuda.current_stream()
    dev = Ks[0].device
    pool = _side_streams.setdefault(dev.index, [])
    while len(pool) < len(Ks) - 1:
        pool.append(torch.cuda.Stream(device=dev))
    order = sorted(range(len(Ks)), key=lambda k: -Ks[k].shape[-1])
    out = [None] * len(Ks)
    used = []
    for slot, k in enumerate(order[1:]):
        s = pool[slot]
        s.wait_stream(cur)                              # K_k was produced on the current stream
        out[k] = _eigh_launch(Ks[k], stream=s)
        used.append(s)
    out[order[0]] = _eigh_launch(Ks[order[0]])
    for s in used:
        cur.wait_stream(s)
    return out


def _eigh_check(infos):
    """One device->host read for all the eigensolves of a call (torch.linalg.eigh raises on failure; so do we)."""
    ops.check_info(torch.cat(infos), 'eigh')


def eigh(K):
    """Ascending eigenpairs of a symmetric matrix (upper triangle read), Jacobi in shared memory.
    No autograd: the Kronecker loss below differentiates analytically w.r.t. K instead of through eigh."""
    w, V, info, _ = _eigh_launch(K)
    _eigh_check([info])
    return w.to(K.dtype), V.to(K.dtype)


def _sizes_arr(sizes):
    return (ctypes.c_int * len(sizes))(*[int(s) for s in sizes])


def _kron_core(T1, lam_cat, sizes, tau_t, add):
    L = B.lib()
    dev = T1.device
    core = torch.empty_like(T1)
    A = torch.empty_like(T1)
    sums = torch.empty(4, dtype=torch.float64, device=dev)
    sb = L.ffgp_kron_core_scratch_bytes(T1.numel())
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    rc = L.ffgp_kron_core_f64(B.ptr(T1), B.ptr(lam_cat), _sizes_arr(sizes), len(sizes), B.ptr(tau_t), float(add),
                              B.ptr(core), B.ptr(A), B.ptr(sums), B.ptr(scratch), sb, B.stream_ptr())
    B.check(rc, 'ffgp_kron_core_f64')
    return core, A, sums


def _kron_scale(inp, lam_cat, sizes, skip, divide_by_A, tau_t, add, dev):
    L = B.lib()
    out = torch.empty([int(s) for s in sizes], dtype=torch.float64, device=dev)
    rc = L.ffgp_kron_scale_f64(B.ptr(inp), B.ptr(lam_cat), _sizes_arr(sizes), len(sizes), int(skip), int(divide_by_A),
                               B.ptr(tau_t), float(add), B.ptr(out), B.stream_ptr())
    B.check(rc, 'ffgp_kron_scale_f64')
    return out


def _kron_ck(lam_cat, sizes, k, tau_t, add, dev):
    """c_k[j] = sum_{i_k = j} prod_{m != k} lambda_m / A: a reduction over the eigenvalues only (ffgp_kron_ck_f64)."""
    L = B.lib()
    nk = int(sizes[k])
    if len(sizes) < 2:                                   # single mode: c_0[j] = 1 / (lambda_0[j] + tau)
        return 1.0 / (lam_cat + tau_t + add)
    out = torch.empty(nk, dtype=torch.float64, device=dev)
    sb = L.ffgp_kron_ck_scratch_bytes(nk)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    rc = L.ffgp_kron_ck_f64(B.ptr(lam_cat), _sizes_arr(sizes), len(sizes), int(k), B.ptr(tau_t), float(add), B.ptr(out),
                            B.ptr(scratch), sb, B.stream_ptr())
    B.check(rc, 'ffgp_kron_ck_f64')
    return out


def _lambda_product(eigvals, sizes, skip=-1):
    """prod_{m != skip} lambda_m[i_m] as a broadcastable tensor (extent 1 along `skip`)."""
    out = None
    for m, lam in enumerate(eigvals):
        if m == skip:
            continue
        shape = [1] * len(sizes)
        shape[m] = int(sizes[m])
        v = lam.reshape(shape)
        out = v if out is None else out * v
    if out is None:
        out = torch.ones([1] * len(sizes), dtype=torch.float64, device=eigvals[0].device)
    return out


class _KronNLL(torch.autograd.Function):
    """S = kron(K_0..K_M) + tau I (+ E);  returns (0.5*log|S| + 0.5*vec(Y)^T S^-1 vec(Y),  A,  g = S^-1 vec(Y)).
    Gradient (closed form, no differentiation through eigh):
      dY = g;  dtau = 0.5*(sum 1/A - sum h^2), h = T1/A;
      dK_k = 0.5 * U_k (diag(c_k) - Q_k) U_k^T,  c_k[j] = sum_{i_k=j} prod_{m!=k} lambda_m / A,
      Q_k = weighted mode-k Gram matrix of h.
    `add` is a number (fused core / reduction kernels) or a tensor E broadcastable to Y's shape, added to A element
    by element in the eigenbasis exactly as the reference's `A = A + y_var` does (hogp.py:176): the eigensolves, mode
    products and mode Grams stay on our kernels, the element-wise stage runs as torch device ops, dE = 0.5 (1/A - h^2).
    With a tensor E the objective is no longer a function of S alone (E lives in the eigenbasis), so the eigenvector
    derivative does not cancel: dK_k = 0.5 U_k (diag(c_k) - Q_k - R_k) U_k^T with
      R_k[a,b] = (B_k[a,b] - B_k[b,a]) / (lambda_a - lambda_b),  B_k = mode-k Gram of (E o h) with h,  R_k[a,a] = 0
    (derived in the docstring's notation from du_a = sum_{b != a} u_b (u_b^T dK u_a) / (lambda_a - lambda_b); it matches
    the reference's autograd through eigh to 1e-13 on tests/golden/hogp2023_yvar.npz).  Exactly equal eigenvalues
    contribute 0 where the reference returns NaN; R_k vanishes when E is constant along mode k."""

    @staticmethod
    def forward(ctx, Y, tau, add, *Ks):
        dev = Y.device
        Yc = ops._f64c(Y)
        sizes = list(Yc.shape)
        E = None
        if isinstance(add, torch.Tensor):
            if add.numel() > 1 or add.requires_grad:      # a one-element tensor that needs its gradient is broadcast too
                E = ops._f64c(add).broadcast_to(sizes)
            else:
                add = float(add)
        launched = _eigh_launch_all(Ks)                   # concurrent per-mode solves, one status read
        _eigh_check([l[2] for l in launched])
        eig = [(l[0], l[1]) for l in launched]
        lam_cat = torch.cat([e[0] for e in eig]).contiguous()
        T1 = Yc
        for k, (_, U) in enumerate(eig):
            T1 = _mode_dot_raw(T1, U.contiguous(), k, True)          # x_k U_k^T
        tau_t = ops._f64c(tau.reshape(1))
        if E is None:
            h, A, sums = _kron_core(T1, lam_cat, sizes, tau_t, add)
            A_saved = E_saved = torch.empty(0, dtype=torch.float64, device=dev)
        else:
            A = (_lambda_product([e[0] for e in eig], sizes) + tau_t + E).contiguous()
            h = (T1 / A).contiguous()
            sums = torch.stack([A.log().sum(), (T1 * h).sum(), (1.0 / A).sum(), (h * h).sum()])
            A_saved, E_saved = A, E.contiguous()
        g = h
        for k, (_, U) in enumerate(eig):
            g = _mode_dot_raw(g, U.contiguous(), k, False)           # x_k U_k
        ctx.save_for_backward(h, g, sums, lam_cat, tau_t, A_saved, E_saved, *[e[1] for e in eig])
        ctx.meta = (sizes, None if E is not None else float(add), Y.dtype, [K.dtype for K in Ks], tau.shape, tau.dtype,
                    (tuple(add.shape), add.dtype) if E is not None else None)
        val = 0.5 * (sums[0] + sums[1])
        A_o, g_o = A.to(Y.dtype), g.to(Y.dtype)
        flat = []
        for lam, U in eig:
            flat += [lam.to(Y.dtype), U.to(Y.dtype)]
        ctx.mark_non_differentiable(A_o, g_o, *flat)
        return (val.to(Y.dtype), A_o, g_o) + tuple(flat)

    @staticmethod
    def backward(ctx, go, *_unused):
        h, g, sums, lam_cat, tau_t, A, E = ctx.saved_tensors[:7]
        Us = ctx.saved_tensors[7:]
        sizes, add, ydt, kdts, tau_shape, tau_dt, e_meta = ctx.meta
        dev = h.device
        go = go.to(torch.float64)
        gY = (g * go).to(ydt) if ctx.needs_input_grad[0] else None
        gtau = None
        if ctx.needs_input_grad[1]:
            gtau = (0.5 * (sums[2] - sums[3]) * go).reshape(tau_shape).to(tau_dt)
        gE = None
        if e_meta is not None and ctx.needs_input_grad[2]:
            gE = (0.5 * go * (1.0 / A - h * h)).sum_to_size(e_meta[0]).to(e_meta[1])
        lams = list(torch.split(lam_cat, [int(n) for n in sizes])) if e_meta is not None else None
        gKs = []
        for k, U in enumerate(Us):
            if not ctx.needs_input_grad[3 + k]:
                gKs.append(None)
                continue
            if e_meta is None:
                c_k = _kron_ck(lam_cat, sizes, k, tau_t, add, dev)                   # sum_{i_k=j} prod_{m!=k} lambda_m / A
                hk = _kron_scale(h, lam_cat, sizes, k, 0, tau_t, add, dev)           # h o prod_{m!=k} lambda_m
            else:
                w = _lambda_product(lams, sizes, skip=k)
                other = [m for m in range(len(sizes)) if m != k]
                c_k = (w / A).sum(dim=other) if other else (w / A).reshape(-1)
                hk = (h * w).contiguous()
            Qk = _mode_gram_raw(h, hk, k)
            Mk = torch.diag(c_k) - Qk
            if e_meta is not None:
                Bk = _mode_gram_raw((E * h).contiguous(), h, k)
                gap = lams[k].reshape(-1, 1) - lams[k].reshape(1, -1)
                nz = gap != 0
                Mk = Mk - torch.where(nz, (Bk - Bk.T) / torch.where(nz, gap, torch.ones_like(gap)), torch.zeros_like(gap))
            Mk = Mk.contiguous()
            Uc = U.contiguous()
            gK = _mode_dot_raw(_mode_dot_raw(Mk, Uc, 0, False), Uc, 1, False)        # U M U^T
            gKs.append((0.5 * go * gK).to(kdts[k]))
        return (gY, gtau, gE) + tuple(gKs)


def kron_nll(Y, Ks, tau, add=0.0):
    """Differentiable Kronecker-GP objective.  Returns (value, A, g, [(lambda_k, U_k)]); value excludes the
    0.5*nd*log(2 pi) constant.  `add`: a number, or a tensor broadcastable to Y (the reference's tensor-valued y_var)."""
    out = _KronNLL.apply(Y, tau, add if isinstance(add, torch.Tensor) else float(add), *Ks)
    val, A, g = out[:3]
    eig = [(out[3 + 2 * k], out[4 + 2 * k]) for k in range(len(Ks))]
    return val, A, g, eig


# `import tensorly; tensorly.tenalg.mode_dot(...)` (hogp.py:132, matrix.py:73): the alias serves both names
tenalg = __import__('sys').modules[__name__]
