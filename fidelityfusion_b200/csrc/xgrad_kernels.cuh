// Gradient of a stationary SE kernel matrix with respect to its INPUT points, and the posterior
// (mean / variance) gradient with respect to the test points built on it.
//
// Reference call chain replaced: the acquisition optimisers differentiate the posterior w.r.t. the candidate
// (MF_BayesianOptimization/Discrete/DMF_acq.py:226-262, Bayesian_optimization v1/MF_EI.py:22-31), i.e. autograd's
// CdistBackward0 / PowBackward / ExpBackward / MmBackward / TriangularSolveBackward chain of cigp.forward
// (GaussianProcess/cigp_v10.py:24-48).  Closed form used here, with W = gK o K:
//     d/d r_k  sum_{r,c} gK[r][c] K(r, c) = -w_k^2 * sum_c W[r][c] (xr[r][k] - xc[c][k])
// ("rows" r are the points differentiated, "cols" c the other argument; K is re-evaluated from the points, never read).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ffgp {

struct XGradParams {
  const double* xr; const double* xc;      // [batch][R][d], [batch][C][d]
  const double* w; long long sw;           // inverse length scales [d] (+ batch stride)
  const double* amp; long long samp;       // [1] (+ batch stride)
  const double* gK;                        // element (r, c) at gK[r * ldr + c * ldc] (+ batch stride sgK); may be NULL
  long long ldr, ldc, sgK;
  double gk_scale;                         // multiplies the gK term (e.g. -1 for the -R (G + G^T) part)
  const double* colscale;                  // optional per-ROW factor on the gK term (diag variance: 2 g_var[r]); [batch][R]
  const double* lr_c; const double* lr_r;  // optional low-rank term  sum_D lr_c[c][D] * lr_r[r][D]  (alpha g_mean^T)
  int lrD; long long ld_lrc, s_lrc, s_lrr;
  int R, C, d, splits;
  long long sxr, sxc;
  double* part;                            // [batch][splits][R][d]
};

constexpr int XG_WARPS = 8;
constexpr int XG_DMAX = 64;

// One warp per row r and column split.  Per 32 columns: lanes = columns compute W[r][c] (distance, exp, weights), then
// lanes = coordinates k accumulate W[r][c] (xr[r][k] - xc[c][k]) with the weights broadcast from shared memory.
__global__ void __launch_bounds__(XG_WARPS * 32) xgrad_kernel(const XGradParams p) {
  __shared__ double wbuf[XG_WARPS][32];
  __shared__ double xrow[XG_WARPS][XG_DMAX];
  __shared__ double wsq[XG_DMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, split = blockIdx.y;
  const int r = blockIdx.x * XG_WARPS + warp;
  const double* wv = p.w + b * p.sw;
  for (int k = threadIdx.x; k < p.d; k += blockDim.x) wsq[k] = wv[k] * wv[k];
  const bool live = r < p.R;
  const double* xr = p.xr + b * p.sxr + (long long)(live ? r : 0) * p.d;
  for (int k = lane; k < p.d; k += 32) xrow[warp][k] = xr[k];
  __syncthreads();
  if (!live) return;
  const double amp = p.amp[b * p.samp];
  const double* xc = p.xc + b * p.sxc;
  const int per = (((p.C + p.splits - 1) / p.splits) + 31) & ~31;
  const int c_lo = split * per, c_hi = min(p.C, c_lo + per);
  const double rs = p.colscale ? p.colscale[(long long)b * p.R + r] : 1.0;
  double acc0 = 0.0, acc1 = 0.0;
  const double xk0 = lane < p.d ? xrow[warp][lane] : 0.0;
  const double xk1 = lane + 32 < p.d ? xrow[warp][lane + 32] : 0.0;
  for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
    const int c = c0 + lane;
    double wt = 0.0;
    if (c < c_hi) {
      const double* xcc = xc + (long long)c * p.d;
      double s = 0.0;
      for (int k = 0; k < p.d; k++) {
        const double df = xrow[warp][k] - xcc[k];
        s = fma(df * df, wsq[k], s);
      }
      double g = 0.0;
      if (p.gK) g = p.gk_scale * rs * p.gK[b * p.sgK + (long long)r * p.ldr + (long long)c * p.ldc];
      if (p.lr_c) {
        const double* a = p.lr_c + b * p.s_lrc + (long long)c * p.ld_lrc;
        const double* gm = p.lr_r + b * p.s_lrr + (long long)r * p.lrD;
        for (int D = 0; D < p.lrD; D++) g = fma(a[D], gm[D], g);
      }
      wt = g * amp * exp(-0.5 * s);
    }
    wbuf[warp][lane] = wt;
    __syncwarp();
    const int m = min(32, c_hi - c0);
    if (lane < p.d) {
      const double* col = xc + (long long)c0 * p.d + lane;
      for (int cc = 0; cc < m; cc++) {
        const double wv2 = wbuf[warp][cc];
        acc0 = fma(wv2, xk0 - col[(long long)cc * p.d], acc0);
        if (lane + 32 < p.d) acc1 = fma(wv2, xk1 - col[(long long)cc * p.d + 32], acc1);
      }
    }
    __syncwarp();
  }
  double* out = p.part + (((long long)b * p.splits + split) * p.R + r) * p.d;
  if (lane < p.d) out[lane] = -wsq[lane] * acc0;
  if (lane + 32 < p.d) out[lane + 32] = -wsq[lane + 32] * acc1;
}

// out[b][r][k] (+)= sum over splits (fixed order) of part[b][s][r][k]
__global__ void xgrad_finish_kernel(const double* __restrict__ part, int splits, long long rd, double* __restrict__ out,
                                    int accumulate) {
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rd) return;
  double s = 0.0;
  for (int sp = 0; sp < splits; sp++) s += part[((long long)b * splits + sp) * rd + i];
  double* o = out + (long long)b * rd + i;
  *o = accumulate ? *o + s : s;
}

// dst (padded [nsp][nsp], zero tail) = G + G^T of the caller's [ns][ns] array
__global__ void sym_pad_kernel(const double* __restrict__ G, int ns, long long sG, double* __restrict__ dst, int nsp,
                               long long sdst) {
  const int b = blockIdx.z;
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= nsp || c >= nsp) return;
  double v = 0.0;
  if (r < ns && c < ns) v = G[b * sG + (long long)r * ns + c] + G[b * sG + (long long)c * ns + r];
  dst[b * sdst + (long long)r * nsp + c] = v;
}

inline int xgrad_splits(int R, int C, int batch) {
  // enough warps to fill 148 SMs x 8 warps x ~2, without cutting a row's columns below 64
  const long long rows = (long long)((R + XG_WARPS - 1) / XG_WARPS) * XG_WARPS * batch;
  long long s = (148LL * XG_WARPS * 2 + rows - 1) / rows;
  const long long smax = (C + 63) / 64;
  if (s > smax) s = smax;
  if (s < 1) s = 1;
  return (int)s;
}

}  // namespace ffgp
