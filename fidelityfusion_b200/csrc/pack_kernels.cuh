// Result row of one batched GP problem, written in ONE launch: chain rule from the kernel-native gradients
// (d/d inv_ls, d/d amp, d/d diag) to the reference's raw parameters (length_scales, signal_variance, log_beta of
// cigp + ARDKernel, GaussianProcess/kernel.py:100-105, cigp_v10.py:57-58), the NLL constant, and the packed layout
// [nll | g_length_scales[d] | g_signal_variance | g_log_beta | mean[ns*D] | var[ns] | info] that the multi-GPU
// all-gather ships.  Replaces ~15 elementwise / cat launches per sweep (part of the 7 % that
// separated 7.44x from 8x at 8 GPUs).  One CTA per problem; fixed-order reduction of the diagonal gradient.
// With acq.kind >= 0 the row also carries the acquisition score of every test point (the erfc / exp
// epilogue of the batched posterior - the candidates' scores leave the sweep in the same row, and through the same
// all-gather, as the predictions they are computed from): [.. | mean[ns] | var[ns] | score[ns] | info], D = 1.
#pragma once
#include <cuda_runtime.h>
#include "acq_kernels.cuh"

namespace ffgp {

struct PackParams {
  const double *nll_core, *g_il, *g_amp, *g_diag, *mean, *var;   // [B], [B][d], [B], [B][n], [B][ns*D], [B][ns]
  const int* info;                                                // [B]
  const double *ls, *sv, *lb;                                     // raw parameters [B][d], [B], [B]
  int B, n, d, D, ns, want_grad, with_info;
  double nll_const, eps;
  double* out; int ld;                                            // [B][ld]
  AcqConsts acq;                                                  // kind < 0: no score block
};

__global__ void __launch_bounds__(128) pack_results_kernel(const PackParams p) {
  const int b = blockIdx.x, tid = threadIdx.x;
  double* o = p.out + (long long)b * p.ld;
  __shared__ double red[128];
  int c = 0;
  if (tid == 0) o[0] = p.nll_core[b] + p.nll_const;
  c = 1;
  if (p.want_grad) {
    double s = 0.0;
    for (int i = tid; i < p.n; i += 128) s += p.g_diag[(long long)b * p.n + i];
    red[tid] = s;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) { if (tid < w) red[tid] += red[tid + w]; __syncthreads(); }
    for (int k = tid; k < p.d; k += 128) {
      const double l = p.ls[(long long)b * p.d + k], ell = fabs(l) + p.eps;
      const double sgn = l > 0.0 ? 1.0 : (l < 0.0 ? -1.0 : 0.0);
      o[c + k] = p.g_il[(long long)b * p.d + k] * (-1.0 / (ell * ell)) * sgn;         // inv_ls = 1 / (|ls| + eps)
    }
    if (tid == 0) {
      const double s_ = p.sv[b];
      o[c + p.d] = p.g_amp[b] * (s_ > 0.0 ? 1.0 : (s_ < 0.0 ? -1.0 : 0.0));           // amp = |sv|
      o[c + p.d + 1] = red[0] * (-exp(-p.lb[b]));                                     // diag = e^-lb + jitter
    }
    c += p.d + 2;
  }
  if (p.ns > 0) {
    const int nm = p.ns * p.D;
    for (int k = tid; k < nm; k += 128) o[c + k] = p.mean[(long long)b * nm + k];
    for (int k = tid; k < p.ns; k += 128) o[c + nm + k] = p.var[(long long)b * p.ns + k];
    c += nm + p.ns;
    if (p.acq.kind >= 0) {                                        // D == 1 (checked by the entry point)
      for (int k = tid; k < p.ns; k += 128) {
        double s, dm, dv;
        acq_score(p.acq, p.mean[(long long)b * p.ns + k], p.var[(long long)b * p.ns + k], s, dm, dv);
        o[c + k] = s;
      }
      c += p.ns;
    }
  }
  if (p.with_info && tid == 0) o[c] = (double)p.info[b];
}

}  // namespace ffgp
