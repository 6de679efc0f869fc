// Acquisition scores on the device: one fused pass over the posterior mean / variance of the
// candidates producing the score AND its partial derivatives, so the candidate optimisers never leave the GPU.
// Reference: MF_BayesianOptimization/Discrete/DMF_acq.py:47-63 (UCB), :82-104 (EI, cdf/pdf through scipy on the
// HOST, rounded to float32 and treated as constants by autograd), :106-128 (PI in its log-density form).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ffgp {

struct AcqParams {
  const double* mean; const double* var;   // [m]
  int m, kind;                             // 0 UCB, 1 EI, 2 PI
  double f_best, beta, xi, std_min, two_pi;
  int round_f32;                           // EI: cdf / pdf rounded to float32 like torch.tensor(norm.cdf(..), dtype=float32)
  double* score; double* d_mean; double* d_var;   // [m]; the derivative outputs may be NULL
};

__global__ void __launch_bounds__(256) acq_kernel(const AcqParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.m) return;
  const double mu = p.mean[i], v = p.var[i];
  double s, dm, dv;
  if (p.kind == 0) {                       // UCB_MF: mean + beta * VARIANCE (DMF_acq.py:62)
    s = mu + p.beta * v; dm = 1.0; dv = p.beta;
  } else {
    const double sd_raw = sqrt(v);
    const bool clamped = !(sd_raw > p.std_min);          // torch.clamp(std, min): gradient 0 below the bound
    const double sd = clamped ? p.std_min : sd_raw;
    const double dsd_dv = clamped ? 0.0 : 0.5 / sd_raw;
    const double t = mu - p.f_best - p.xi;
    const double z = t / sd;
    if (p.kind == 1) {                     // EI_MF
      double cdf = 0.5 * erfc(-z * 0.70710678118654752440);
      double pdf = exp(-0.5 * z * z) * 0.39894228040143267794;
      if (p.round_f32) { cdf = (double)__double2float_rn(cdf); pdf = (double)__double2float_rn(pdf); }
      s = t * cdf + sd * pdf; dm = cdf; dv = pdf * dsd_dv;
    } else {                               // PI_MF: -Z^2/2 - log(1) - log(sqrt(2 PI))
      s = -0.5 * z * z - 0.5 * log(p.two_pi);
      dm = -z / sd; dv = (z * z / sd) * dsd_dv;
    }
  }
  p.score[i] = s;
  if (p.d_mean) p.d_mean[i] = dm;
  if (p.d_var) p.d_var[i] = dv;
}

}  // namespace ffgp
