// Phase timeline of factor256_kernel (clock64 stamps at the phase boundaries, one CTA).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFFGP_F256_TRACE -o tools/f256_trace tools/f256_trace.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../fidelityfusion_b200/csrc/factor256.cuh"
using namespace ffgp;
int main() {
  const int n = 256;
  std::vector<double> h(n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) h[i * n + j] = exp(-0.05 * (i - j) * (i - j)) + (i == j ? 0.5 : 0.0);
  double *A, *L, *M, *ld; int* info;
  cudaMalloc(&A, n * n * 8); cudaMalloc(&L, n * n * 8); cudaMalloc(&M, n * n * 8); cudaMalloc(&ld, 64); cudaMalloc(&info, 4);
  cudaMemcpy(A, h.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemset(info, 0, 4);
  cudaMemset(L, 0, n * n * 8); cudaMemset(M, 0, n * n * 8);
  cudaFuncSetAttribute(factor256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F2_SMEM);
  for (int skip = 0; skip < 2; skip++) {
    for (int it = 0; it < 3; it++) factor256_kernel<<<1, 256, F2_SMEM>>>(A, L, M, n, 0, ld, 4, 0, info, 0, skip);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; it++) factor256_kernel<<<1, 256, F2_SMEM>>>(A, L, M, n, 0, ld, 4, 0, info, 0, skip);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("factor256_kernel skip_tri=%d: %.2f us per launch (%s)\n", skip, ms * 100, cudaGetErrorString(cudaGetLastError()));
    long long tr[16];
    cudaMemcpyFromSymbol(tr, g_f256_trace, sizeof(tr));
    const char* names[9] = {"load A11", "base(A11)", "store L11 M11", "P2 L21 = A21 M11^T (+store, park)", "P3 A22 -= L21 L21^T (+unpark)",
                            "base(A22)", "store L22 M22", "P4 T = M22 L21 (+park)", "P5 M21 = -T M11 (+store)"};
    for (int k = 0; k < 9; k++) printf("  %-36s %7lld clk\n", names[k], tr[k + 1] - tr[k]);
    printf("  %-36s %7lld clk\n", "total", tr[9] - tr[0]);
    long long t2[8 * 8 * 3];
    cudaMemcpyFromSymbol(t2, g_f256_trace2, sizeof(t2));
    printf("  P5 per k16 step, (wait+barrier | k-step) clk for warps 0, 3, 4, 7:\n");
    for (int kt = 0; kt < 8; kt++) {
      printf("   kt %d:", kt);
      for (int w : {0, 3, 4, 7}) printf("  w%d %5lld | %5lld", w, t2[(w * 8 + kt) * 3 + 1] - t2[(w * 8 + kt) * 3 + 0], t2[(w * 8 + kt) * 3 + 2] - t2[(w * 8 + kt) * 3 + 1]);
      printf("\n");
    }
  }
#ifdef FFGP_BASE_TRACE
  {  // stamps of the SECOND base_factor_smem call (it overwrites the first one's)
    std::vector<long long> tr(8 * 16 * 4);
    cudaMemcpyFromSymbol(tr.data(), g_base_trace, tr.size() * 8);
    long long t0 = tr[0];
    printf("base(A22) panel: t(top)  | w0: solve  diag-upd  chol8+publish | w1: solve+wait  hoist  update | panel total\n");
    for (int p = 0; p < 15; p++) {
      auto w0 = [&](int s) { return tr[(0 * 16 + p) * 4 + s]; };
      auto w1 = [&](int s) { return tr[(1 * 16 + p) * 4 + s]; };
      long long nxt = tr[(0 * 16 + p + 1) * 4 + 0];
      printf("%2d: t=%7lld | %5lld %5lld %5lld | %5lld %5lld %5lld | %6lld\n", p, w0(0) - t0, w0(1) - w0(0), w0(2) - w0(1),
             w0(3) - w0(2), w1(1) - w1(0), w1(2) - w1(1), w1(3) - w1(2), nxt - w0(0));
    }
  }
#endif
  std::vector<double> hl(n * n), hm(n * n);
  cudaMemcpy(hl.data(), L, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hm.data(), M, n * n * 8, cudaMemcpyDeviceToHost);
  double err1 = 0, err2 = 0;
  for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
    double s1 = 0, s2 = 0;
    for (int k = 0; k <= j; k++) s1 += hl[i * n + k] * hl[j * n + k];
    for (int k = j; k <= i; k++) s2 += hm[i * n + k] * hl[k * n + j];
    err1 = fmax(err1, fabs(s1 - h[i * n + j])); err2 = fmax(err2, fabs(s2 - (i == j ? 1.0 : 0.0)));
  }
  printf("max |L L^T - A| = %.3e   max |M L - I| = %.3e\n", err1, err2);
  return 0;
}
