// Acquisition scores on the device: one fused pass over the posterior mean / variance of the
// candidates producing the score AND its partial derivatives, so the candidate optimisers never leave the GPU.
// Reference: MF_BayesianOptimization/Discrete/DMF_acq.py:47-63 (UCB), :82-104 (EI, cdf/pdf through scipy on the
// HOST, rounded to float32 and treated as constants by autograd), :106-128 (PI in its log-density form).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ffgp {

// Score and partials of ONE candidate.  kinds 0-2: DiscreteAcquisitionFunction (DMF_acq.py); kinds 3-4: the
// single-fidelity classes of Bayesian_optimization/acq.py (UCB :135-149 on the standard deviation, PI :211-231 as the
// float32-rounded normal cdf, a constant for autograd like EI's cdf / pdf).
struct AcqConsts { int kind; double f_best, beta, xi, std_min, two_pi; int round_f32; };

__device__ __forceinline__ void acq_score(const AcqConsts& c, const double mu, const double v, double& s, double& dm, double& dv) {
  if (c.kind == 0) {                       // UCB_MF: mean + beta * VARIANCE (DMF_acq.py:62)
    s = mu + c.beta * v; dm = 1.0; dv = c.beta;
    return;
  }
  const double sd_raw = sqrt(v);
  if (c.kind == 3) {                       // UCB: mean + kappa * sqrt(variance) (Bayesian_optimization/acq.py:147-149)
    s = mu + c.beta * sd_raw; dm = 1.0; dv = c.beta * 0.5 / sd_raw;
    return;
  }
  const bool clamped = !(sd_raw > c.std_min);            // torch.clamp(std, min): gradient 0 below the bound
  const double sd = clamped ? c.std_min : sd_raw;
  const double dsd_dv = clamped ? 0.0 : 0.5 / sd_raw;
  const double t = mu - c.f_best - c.xi;
  const double z = t / sd;
  if (c.kind == 1) {                       // EI_MF / EI
    double cdf = 0.5 * erfc(-z * 0.70710678118654752440);
    double pdf = exp(-0.5 * z * z) * 0.39894228040143267794;
    if (c.round_f32) { cdf = (double)__double2float_rn(cdf); pdf = (double)__double2float_rn(pdf); }
    s = t * cdf + sd * pdf; dm = cdf; dv = pdf * dsd_dv;
  } else if (c.kind == 2) {                // PI_MF: -Z^2/2 - log(1) - log(sqrt(2 PI))
    s = -0.5 * z * z - 0.5 * log(c.two_pi);
    dm = -z / sd; dv = (z * z / sd) * dsd_dv;
  } else {                                 // kind 4, PI: Phi(Z) built from a numpy array - no gradient (acq.py:230)
    double cdf = 0.5 * erfc(-z * 0.70710678118654752440);
    if (c.round_f32) cdf = (double)__double2float_rn(cdf);
    s = cdf; dm = 0.0; dv = 0.0;
  }
}

struct AcqParams {
  const double* mean; const double* var;   // [m]
  int m;
  AcqConsts c;
  double* score; double* d_mean; double* d_var;   // [m]; the derivative outputs may be NULL
};

__global__ void __launch_bounds__(256) acq_kernel(const AcqParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.m) return;
  double s, dm, dv;
  acq_score(p.c, p.mean[i], p.var[i], s, dm, dv);
  p.score[i] = s;
  if (p.d_mean) p.d_mean[i] = dm;
  if (p.d_var) p.d_var[i] = dv;
}

}  // namespace ffgp
