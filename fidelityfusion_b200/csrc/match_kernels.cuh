// Row matching between two fidelities' input sets: match[i] = smallest j with b[j][:] == a[i][:]
// (IEEE equality, like torch's `==`: NaN matches nothing, -0.0 == 0.0), or -1.  The reference builds the full
// [na][nb][d] boolean tensor (FidelityFusion_Models/MF_data.py:199-202, 235-238: O(na nb d) BYTES); here b streams
// through shared memory in tiles and every thread keeps one row of a, so the only traffic is the two inputs.
#pragma once
#include <cuda_runtime.h>

namespace ffgp {

constexpr int MATCH_ROWS = 128;     // rows of a per CTA = rows of b per tile

__global__ void __launch_bounds__(MATCH_ROWS) row_match_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                               int na, int nb, int d, int* __restrict__ match) {
  extern __shared__ double msm[];
  const int lda = d + 1;                                  // odd-ish stride: a thread walks its own row conflict-free
  double* as = msm;                                       // [MATCH_ROWS][d + 1]
  double* bs = msm + MATCH_ROWS * lda;                    // [MATCH_ROWS][d]   (read as a broadcast)
  const int tid = threadIdx.x, i0 = blockIdx.x * MATCH_ROWS, i = i0 + tid;
  for (int e = tid; e < MATCH_ROWS * d; e += MATCH_ROWS) {
    const int r = e / d, k = e - r * d;
    as[r * lda + k] = (i0 + r < na) ? a[(long long)(i0 + r) * d + k] : 0.0;
  }
  int found = -1;
  for (int j0 = 0; j0 < nb; j0 += MATCH_ROWS) {
    __syncthreads();
    const int nj = min(MATCH_ROWS, nb - j0);
    for (int e = tid; e < nj * d; e += MATCH_ROWS) bs[e] = b[(long long)j0 * d + e];
    __syncthreads();
    if (found < 0 && i < na) {
      for (int j = 0; j < nj; j++) {
        bool eq = true;
        for (int k = 0; k < d && eq; k++) eq = (as[tid * lda + k] == bs[j * d + k]);
        if (eq) { found = j0 + j; break; }
      }
    }
  }
  if (i < na) match[i] = found;
}

}  // namespace ffgp
