// Internal timeline of potrf_trtri_base_kernel (clock64 stamps per warp / panel / phase).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFFGP_BASE_TRACE [-DTRACE_BN=64] -o tools/base_trace tools/base_trace.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../fidelityfusion_b200/csrc/dense_kernels.cuh"
using namespace ffgp;
#ifndef TRACE_BN
#define TRACE_BN 128
#endif
int main() {
  const int n = TRACE_BN;
  std::vector<double> h(n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) h[i * n + j] = exp(-0.05 * (i - j) * (i - j)) + (i == j ? 0.5 : 0.0);
  double *A, *L, *M, *ld; int* info;
  cudaMalloc(&A, n * n * 8); cudaMalloc(&L, n * n * 8); cudaMalloc(&M, n * n * 8); cudaMalloc(&ld, 64); cudaMalloc(&info, 4);
  cudaMemcpy(A, h.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemset(info, 0, 4);
  cudaFuncSetAttribute(potrf_trtri_base_kernel<TRACE_BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base_smem_bytes(TRACE_BN));
  for (int it = 0; it < 3; it++) potrf_trtri_base_kernel<TRACE_BN><<<1, 256, base_smem_bytes(TRACE_BN)>>>(A, L, M, n, 0, ld, 1, 0, info, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int it = 0; it < 10; it++) potrf_trtri_base_kernel<TRACE_BN><<<1, 256, base_smem_bytes(TRACE_BN)>>>(A, L, M, n, 0, ld, 1, 0, info, 0);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("base kernel: %.2f us per launch (%s)\n", ms * 100, cudaGetErrorString(cudaGetLastError()));
  std::vector<long long> tr(8 * 16 * 4);
  cudaMemcpyFromSymbol(tr.data(), g_base_trace, tr.size() * 8);
  long long t0 = tr[0];
  printf("panel: t(top)  | w0: solve  diag-upd  chol8+publish | w1: solve+wait  hoist  update | panel total\n");
  for (int p = 0; p < TRACE_BN / 8 - 1; p++) {
    auto w0 = [&](int s) { return tr[(0 * 16 + p) * 4 + s]; };
    auto w1 = [&](int s) { return tr[(1 * 16 + p) * 4 + s]; };
    long long nxt = tr[(0 * 16 + p + 1) * 4 + 0];
    printf("%2d: t=%7lld | %5lld %5lld %5lld | %5lld %5lld %5lld | %6lld\n", p, w0(0) - t0, w0(1) - w0(0), w0(2) - w0(1),
           w0(3) - w0(2), w1(1) - w1(0), w1(2) - w1(1), w1(3) - w1(2), nxt - w0(0));
  }
  // correctness: L L^T = A and M L = I
  std::vector<double> hl(n * n), hm(n * n);
  cudaMemcpy(hl.data(), L, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hm.data(), M, n * n * 8, cudaMemcpyDeviceToHost);
  double err1 = 0, err2 = 0;
  for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
    double s1 = 0, s2 = 0;
    for (int k = 0; k < n; k++) { s1 += hl[i * n + k] * hl[j * n + k]; s2 += hm[i * n + k] * hl[k * n + j]; }
    err1 = fmax(err1, fabs(s1 - h[i * n + j])); err2 = fmax(err2, fabs(s2 - (i == j ? 1.0 : 0.0)));
  }
  printf("max |L L^T - A| = %.3e   max |M L - I| = %.3e\n", err1, err2);
  return 0;
}
