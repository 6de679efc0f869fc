// Device-side training step: the optimiser update fused into one launch that follows the gradient
// kernels, with the step counter and the loss history kept on the device so a captured CUDA graph of one epoch can be
// replayed without any host value changing between iterations.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ffgp {

// torch.optim.Adam (no weight decay, no amsgrad), operation order of torch/optim/adam.py::_single_tensor_adam:
//   m <- lerp(m, g, 1 - b1);  v <- b2 v + (1 - b2) g g;  p <- p - (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// table[5 i + {0..4}] = {param, grad, exp_avg, exp_avg_sq, step} of tensor i (sizes[i] elements; step is a device
// double[1] per tensor, torch keeps the step count per parameter too); one CTA per tensor, which also advances that
// tensor's step.  loss_hist[t-1] = loss[0] for the first hist_cap steps (t of tensor 0).
__global__ void __launch_bounds__(256) adam_step_kernel(void* const* __restrict__ table, const int* __restrict__ sizes,
                                                        double lr, double b1, double b2, double eps, int maximize,
                                                        const double* __restrict__ loss, double* __restrict__ loss_hist,
                                                        int hist_cap) {
  const int i = blockIdx.x;
  double* prm = static_cast<double*>(table[5 * i + 0]);
  const double* grd = static_cast<const double*>(table[5 * i + 1]);
  double* m = static_cast<double*>(table[5 * i + 2]);
  double* v = static_cast<double*>(table[5 * i + 3]);
  double* step = static_cast<double*>(table[5 * i + 4]);
  const double t = step[0] + 1.0;
  const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  const double step_size = lr / bc1, bc2_sqrt = sqrt(bc2);
  const int n = sizes[i];
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const double g = maximize ? -grd[e] : grd[e];
    const double mo = m[e];
    const double mn = mo + (1.0 - b1) * (g - mo);
    const double vn = v[e] * b2 + (1.0 - b2) * g * g;
    m[e] = mn;
    v[e] = vn;
    prm[e] = prm[e] - step_size * (mn / (sqrt(vn) / bc2_sqrt + eps));
  }
  __syncthreads();                                  // every thread has read the old step count
  if (threadIdx.x == 0) {
    step[0] = t;
    if (i == 0 && loss && loss_hist && t <= (double)hist_cap) loss_hist[(int)t - 1] = loss[0];
  }
}

}  // namespace ffgp
