"""High-precision arbiter for the HOGP hyper-parameter gradients (tests/golden/hogp2023_arbiter.npz).

TEST INFRASTRUCTURE.  The reference differentiates its Kronecker loss THROUGH torch.linalg.eigh (hogp.py:18-22,
140-198); the CUDA path uses the closed form.  Where the two disagree beyond 1e-9 somebody is wrong, and fp64 finite
differences cannot tell (eps / h).  This script evaluates the loss of the `hogp2023_params` case
(oracle/gen_golden.py:gen_c4) in 50-digit arithmetic (mpmath: SE kernel -> symmetric eigen-decomposition -> A, T1 ->
loss) and differentiates it by central differences with h = 1e-12: truncation ~1e-24, rounding ~1e-38.

    python oracle/gen_golden_arbiter.py        (about a minute)"""
import os

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, '..', 'tests', 'golden')
mp.mp.dps = 50


def se_kernel(X, ls, sc):
    """MFGP_ver2023May/kernel/SE_kernel.py:20-44, linear format (what kernel_utils.create_kernel builds):
    K = sc * exp(-0.5 * (|x/ls|^2 + |x'/ls|^2 - 2 x.x'/ls^2))."""
    n = len(X)
    Z = [[v / ls for v in row] for row in X]
    sq = [sum(v * v for v in row) for row in Z]
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            dot = sum(a * b for a, b in zip(Z[i], Z[j]))
            K[i, j] = sc * mp.exp(-(sq[i] + sq[j] - 2 * dot) / 2)
    return K


def mode_dot_T(T, U, mode, shape):
    """T x_mode U^T for a flat row-major tensor T of `shape` (U: [n, n] with eigenvectors as columns)."""
    n = shape[mode]
    outer = int(np.prod(shape[:mode]))
    inner = int(np.prod(shape[mode + 1:]))
    out = [mp.mpf(0)] * len(T)
    for o in range(outer):
        for i in range(inner):
            col = [T[(o * n + a) * inner + i] for a in range(n)]
            for b in range(n):
                out[(o * n + b) * inner + i] = sum(U[a, b] * col[a] for a in range(n))
    return out


def loss(x, Y, shape, ls, sc, noise):
    ins = [x] + [[[mp.mpf(i)] for i in range(s)] for s in shape[1:]]
    lam, T = [], Y
    for k in range(len(shape)):
        E, Q = mp.eigsy(se_kernel(ins[k], ls[k], sc[k]))
        lam.append([E[i] for i in range(shape[k])])
        T = mode_dot_T(T, Q, k, shape)
    nd = len(Y)
    quad, logdet = mp.mpf(0), mp.mpf(0)
    idx = [0] * len(shape)
    for e in range(nd):
        r = e
        for k in range(len(shape) - 1, -1, -1):
            idx[k] = r % shape[k]
            r //= shape[k]
        a = 1 / noise
        p = mp.mpf(1)
        for k in range(len(shape)):
            p *= lam[k][idx[k]]
        a += p
        logdet += mp.log(a)
        quad += T[e] * T[e] / a
    return (nd * mp.log(2 * mp.pi) / 2 + logdet / 2 + quad / 2) / nd


def main():
    with np.load(os.path.join(GOLD, 'hogp2023_params.npz')) as z:
        x = [[mp.mpf(float(v)) for v in row] for row in z['x']]
        Yn = z['Y']
    shape = list(Yn.shape)
    Y = [mp.mpf(float(v)) for v in Yn.reshape(-1)]
    ls0 = [mp.mpf(2 * (i + 1)) / 10 - mp.mpf(3) / 10 for i in range(4)]
    # the fp64 parameters the models hold are the fp64 roundings of these decimals: use exactly those
    ls0 = [mp.mpf(float(0.2 * (i + 1) - 0.3)) for i in range(4)]
    sc0 = [mp.mpf(float(0.1 * (i + 1))) for i in range(4)]
    nz0 = mp.mpf(3)
    h = mp.mpf(10) ** -12
    out = {'loss': float(loss(x, Y, shape, ls0, sc0, nz0))}
    for k in range(4):
        for name, base in (('length_scale', ls0), ('scale', sc0)):
            up, dn = list(base), list(base)
            up[k] += h
            dn[k] -= h
            if name == 'length_scale':
                d = (loss(x, Y, shape, up, sc0, nz0) - loss(x, Y, shape, dn, sc0, nz0)) / (2 * h)
            else:
                d = (loss(x, Y, shape, ls0, up, nz0) - loss(x, Y, shape, ls0, dn, nz0)) / (2 * h)
            out[f'g_kernel_list_{k}_{name}'] = float(d)
            print(k, name, mp.nstr(d, 20), flush=True)
    d = (loss(x, Y, shape, ls0, sc0, nz0 + h) - loss(x, Y, shape, ls0, sc0, nz0 - h)) / (2 * h)
    out['g_noise_box_value'] = float(d)
    np.savez(os.path.join(GOLD, 'hogp2023_arbiter.npz'), **{k: np.asarray(v) for k, v in out.items()})
    print('wrote hogp2023_arbiter.npz', out)


if __name__ == '__main__':
    main()
