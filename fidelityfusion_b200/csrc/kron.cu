// Kronecker / Tucker path of the GP hot path and the stand-alone
// kernel-matrix backward:
//   mode_dot_kernel     n-mode product  out = t x_mode mat  on the FP64 tensor pipe (DMMA), the small
//                       factor matrix resident in shared memory, the tensor streamed once
//   mode_gram_kernel    G[a][b] = sum_rest X[a,rest] Y[b,rest] along a mode (gradient of mode_dot w.r.t. its
//                       matrix, and the weighted Gram matrices of the Kronecker-GP gradient), split-K + fixed-order reduce
//   kron_core_kernel    A = kron(lambda) + tau, core = T1/A, sum log A, sum T1^2/A, ... in one pass
//   kron_scale_kernel   out = in o prod_{m != skip} lambda_m
//   syevj_kernel        cyclic parallel-order Jacobi eigensolver, one CTA per matrix
//   kernel_bwd_kernel   d(sum gK o K)/d(inv_ls, amp) for a rectangular kernel matrix
#include <algorithm>
#include <cstring>
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/ffgp.h"
#include "gemm_dmma.cuh"

namespace ffgp {
int fail(int code, const char* fmt, const char* a);
extern unsigned long long g_launches;

#define FFGP_CUDA(x)                                                   \
  do {                                                                 \
    cudaError_t e__ = (x);                                             \
    if (e__ != cudaSuccess) return fail(-100, "CUDA error: %s", cudaGetErrorString(e__)); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Generalised columns: a contiguous tensor viewed as [outer][I][inner] is a K x Ncols operand with
//   element(k, c) at  (c / inner) * I * inner + k * inner + (c % inner),  Ncols = outer * inner.
// This covers every mode (inner == 1 is the last mode) with one addressing rule.
// out(j, c) = sum_k mat(j, k) * t(k, c);  mat is [J][I] row-major, or [I][J] when transposed.
// CTA: up to 128 rows of J (grid.y row blocks), 64 columns per step, grid-stride over column chunks.
// ---------------------------------------------------------------------------------------------
constexpr int MD_COLS = 64;
constexpr int MD_KC = 16;

__global__ void __launch_bounds__(256) mode_dot_kernel(const double* __restrict__ t, const double* __restrict__ mat,
                                                       double* __restrict__ out, long long ncols, int I, long long inner,
                                                       int J, int transpose_mat) {
  __shared__ double ms[128][MD_KC + 4];          // mat block rows x k-chunk
  __shared__ double bs[MD_KC][MD_COLS + 4];      // tensor chunk  k x cols
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int j0 = blockIdx.y * 128;
  const int jrows = min(128, J - j0);
  const int mtiles = (jrows + 7) / 8;
  const long long nchunks = (ncols + MD_COLS - 1) / MD_COLS;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long c0 = ch * MD_COLS;
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    for (int k0 = 0; k0 < I; k0 += MD_KC) {
      __syncthreads();
      for (int e = tid; e < 128 * MD_KC; e += 256) {
        const int r = e / MD_KC, k = e % MD_KC;
        double v = 0.0;
        if (r < jrows && k0 + k < I)
          v = transpose_mat ? mat[(long long)(k0 + k) * J + j0 + r] : mat[(long long)(j0 + r) * I + k0 + k];
        ms[r][k] = v;
      }
      if (inner == 1) {      // k is the contiguous index
        for (int e = tid; e < MD_KC * MD_COLS; e += 256) {
          const int c = e / MD_KC, k = e % MD_KC;
          double v = 0.0;
          if (c0 + c < ncols && k0 + k < I) v = t[(c0 + c) * I + k0 + k];
          bs[k][c] = v;
        }
      } else {
        for (int e = tid; e < MD_KC * MD_COLS; e += 256) {
          const int k = e / MD_COLS, c = e % MD_COLS;
          double v = 0.0;
          const long long cc = c0 + c;
          if (cc < ncols && k0 + k < I) v = t[(cc / inner) * I * inner + (long long)(k0 + k) * inner + (cc % inner)];
          bs[k][c] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < MD_KC / 4; kk++) {
        const double bf = bs[kk * 4 + tq][warp * 8 + g];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (i < mtiles) {
            const double af = ms[i * 8 + g][kk * 4 + tq];
            dmma884(acc[i][0], acc[i][1], af, bf);
          }
        }
      }
    }
    // C fragment: row g of m-tile i, columns warp*8 + 2*tq + {0,1}
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (i < mtiles) {
        const int r = i * 8 + g;
        if (r < jrows) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const long long cc = c0 + warp * 8 + tq * 2 + e;
            if (cc < ncols) out[(cc / inner) * J * inner + (long long)(j0 + r) * inner + (cc % inner)] = acc[i][e];
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Mode Gram:  G[a][b] = sum_c X(a, c) * Y(b, c)  over generalised columns, X viewed [outer][Ja][inner],
// Y viewed [outer][Jb][inner].  CTA (bx, by, bz): 64x64 block (by, bx) of G over the bz-th slice of columns;
// slices are summed in fixed order by gram_reduce_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int MG_T = 64;
constexpr int MG_KC = 16;

__global__ void __launch_bounds__(256) mode_gram_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                                                        double* __restrict__ part, long long ncols, long long inner,
                                                        int Ja, int Jb, long long cols_per_slice) {
  __shared__ double xs[MG_T][MG_KC + 4];
  __shared__ double ys[MG_T][MG_KC + 4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;              // 2 x 4 warps: warp tile 32 x 16
  const int a0 = blockIdx.y * MG_T, b0 = blockIdx.x * MG_T;
  const long long cbeg = (long long)blockIdx.z * cols_per_slice;
  const long long cend = min(ncols, cbeg + cols_per_slice);
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  for (long long c0 = cbeg; c0 < cend; c0 += MG_KC) {
    __syncthreads();
    for (int e = tid; e < 2 * MG_T * MG_KC; e += 256) {
      const int which = e / (MG_T * MG_KC);
      int r, k;
      if (inner == 1) { r = e % MG_T; k = (e / MG_T) % MG_KC; }    // rows are contiguous in memory for a fixed column
      else { k = e % MG_KC; r = (e / MG_KC) % MG_T; }              // columns (inner index) are contiguous
      const long long cc = c0 + k;
      const int J = which ? Jb : Ja, row = (which ? b0 : a0) + r;
      double v = 0.0;
      if (cc < cend && row < J) {
        const double* src = which ? Y : X;
        v = src[(cc / inner) * J * inner + (long long)row * inner + (cc % inner)];
      }
      (which ? ys : xs)[r][k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < MG_KC / 4; kk++) {
      double af[4], bf[2];
#pragma unroll
      for (int i = 0; i < 4; i++) af[i] = xs[wm * 32 + i * 8 + g][kk * 4 + tq];
#pragma unroll
      for (int j = 0; j < 2; j++) bf[j] = ys[wn * 16 + j * 8 + g][kk * 4 + tq];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  double* dst = part + (long long)blockIdx.z * Ja * Jb;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int a = a0 + wm * 32 + i * 8 + g;
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int b = b0 + wn * 16 + j * 8 + tq * 2 + e;
        if (a < Ja && b < Jb) dst[(long long)a * Jb + b] = acc[i][j][e];
      }
  }
}

__global__ void gram_reduce_kernel(const double* __restrict__ part, int nslices, long long nelem, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelem) return;
  double s = 0.0;
  for (int z = 0; z < nslices; z++) s += part[(long long)z * nelem + i];
  out[i] = s;
}

// ---------------------------------------------------------------------------------------------
// Kronecker core stage
// ---------------------------------------------------------------------------------------------
struct KronSizes { int n[8]; int off[8]; int nmodes; };

__device__ __forceinline__ double kron_lambda_prod(long long idx, const KronSizes& s, const double* __restrict__ lam, int skip) {
  double p = 1.0;
#pragma unroll
  for (int m = 7; m >= 0; m--) {
    if (m < s.nmodes) {
      const int im = (int)(idx % s.n[m]);
      idx /= s.n[m];
      if (m != skip) p *= lam[s.off[m] + im];
    }
  }
  return p;
}

__global__ void __launch_bounds__(256) kron_core_kernel(const double* __restrict__ T1, const double* __restrict__ lam,
                                                        KronSizes s, const double* __restrict__ noise_inv, double add,
                                                        long long total, double* __restrict__ core, double* __restrict__ Aout,
                                                        double* __restrict__ part) {
  __shared__ double red[4][256];
  const double tau = (noise_inv ? noise_inv[0] : 0.0) + add;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const double A = kron_lambda_prod(i, s, lam, -1) + tau;
    const double t1 = T1[i];
    const double ia = 1.0 / A;
    const double h = t1 * ia;
    if (core) core[i] = h;
    if (Aout) Aout[i] = A;
    s0 += log(A);
    s1 = fma(t1, h, s1);
    s2 += ia;
    s3 = fma(h, h, s3);
  }
  red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = s2; red[3][threadIdx.x] = s3;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int q = 0; q < 4; q++) red[q][threadIdx.x] += red[q][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 4) part[(long long)blockIdx.x * 4 + threadIdx.x] = red[threadIdx.x][0];
}

__global__ void kron_sums_finish_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ out) {
  const int q = threadIdx.x;
  if (q >= 4) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; b++) s += part[(long long)b * 4 + q];
  out[q] = s;
}

__global__ void kron_scale_kernel(const double* __restrict__ in, const double* __restrict__ lam, KronSizes s, int skip,
                                  int power_inv_A, const double* __restrict__ noise_inv, double add, long long total,
                                  double* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    double v = (in ? in[i] : 1.0) * kron_lambda_prod(i, s, lam, skip);
    if (power_inv_A) v /= (kron_lambda_prod(i, s, lam, -1) + (noise_inv ? noise_inv[0] : 0.0) + add);
    out[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Jacobi eigensolver (cyclic, round-robin parallel ordering).  One CTA per matrix.
//   a: n x lda working copy of A (shared memory when n <= 128, else global), vt: rows = eigenvectors
// ---------------------------------------------------------------------------------------------
constexpr int SYEVJ_THREADS = 512;
constexpr int SYEVJ_SMEM_N = 128;

__device__ __forceinline__ void rr_pair(int step, int k, int m, int& p, int& q) {
  const int mm = m - 1;
  int a, b;
  if (k == 0) { a = step % mm; b = mm; }
  else { a = (step + k) % mm; b = (step - k + mm) % mm; }
  p = min(a, b); q = max(a, b);
}

__global__ void __launch_bounds__(SYEVJ_THREADS) syevj_kernel(const double* __restrict__ Ain, int n, double* __restrict__ w,
                                                              double* __restrict__ V, double* __restrict__ work_a,
                                                              double* __restrict__ work_vt, int* __restrict__ info,
                                                              int max_sweeps) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int m = (n + 1) & ~1;
  const int half = m / 2;
  const bool in_smem = (n <= SYEVJ_SMEM_N);
  const int lda = in_smem ? (n + 1) : n;
  double* a = in_smem ? sm : work_a + (long long)b * n * n;
  double* rc = in_smem ? sm + (size_t)n * lda : sm;          // [half] cos
  double* rs = rc + half;                                     // [half] sin
  double* red = rs + half;                                    // [nt] reduction scratch
  double* vt = work_vt + (long long)b * n * n;
  const double* A0 = Ain + (long long)b * n * n;
  // load: symmetrise from the UPPER triangle (torch.linalg.eigh(K, UPLO='U'))
  double nrm_local = 0.0;
  for (int e = tid; e < n * n; e += nt) {
    const int i = e / n, j = e % n;
    const double v = (j >= i) ? A0[(long long)i * n + j] : A0[(long long)j * n + i];
    a[i * lda + j] = v;
    vt[e] = (i == j) ? 1.0 : 0.0;
    nrm_local = fma(v, v, nrm_local);
  }
  red[tid] = nrm_local;
  __syncthreads();
  for (int o = nt / 2; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  const double norm2 = red[0];
  __syncthreads();
  int converged = (norm2 == 0.0);
  for (int sweep = 0; sweep < max_sweeps && !converged; sweep++) {
    double off_local = 0.0;
    for (int step = 0; step < m - 1; step++) {
      for (int k = tid; k < half; k += nt) {
        int p, q; rr_pair(step, k, m, p, q);
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = a[p * lda + q], app = a[p * lda + p], aqq = a[q * lda + q];
          off_local = fma(apq, apq, off_local);
          if (fabs(apq) > 1e-300 && fabs(apq) > 1e-19 * (fabs(app) + fabs(aqq))) {
            const double tau = (aqq - app) / (2.0 * apq);
            const double tt = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + tt * tt);
            s = tt * c;
          }
        }
        rc[k] = c; rs[k] = s;
      }
      __syncthreads();
      // rows p,q of A and of V^T
      for (int e = tid; e < half * n; e += nt) {
        const int k = e / n, j = e % n;
        const double s = rs[k];
        if (s != 0.0) {
          const double c = rc[k];
          int p, q; rr_pair(step, k, m, p, q);
          const double ap = a[p * lda + j], aq = a[q * lda + j];
          a[p * lda + j] = c * ap - s * aq;
          a[q * lda + j] = s * ap + c * aq;
          const double vp = vt[(long long)p * n + j], vq = vt[(long long)q * n + j];
          vt[(long long)p * n + j] = c * vp - s * vq;
          vt[(long long)q * n + j] = s * vp + c * vq;
        }
      }
      __syncthreads();
      // columns p,q of A
      for (int e = tid; e < half * n; e += nt) {
        const int i = e / half, k = e % half;
        const double s = rs[k];
        if (s != 0.0) {
          const double c = rc[k];
          int p, q; rr_pair(step, k, m, p, q);
          const double ap = a[i * lda + p], aq = a[i * lda + q];
          a[i * lda + p] = c * ap - s * aq;
          a[i * lda + q] = s * ap + c * aq;
        }
      }
      __syncthreads();
    }
    red[tid] = off_local;
    __syncthreads();
    for (int o = nt / 2; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
    const double off2 = red[0];
    __syncthreads();
    if (off2 <= 1e-26 * norm2) converged = 1;
  }
  if (tid == 0 && !converged) info[b] = 1;
  // ascending sort by rank, ties broken by index
  for (int i = tid; i < n; i += nt) {
    const double li = a[i * lda + i];
    int rank = 0;
    for (int j = 0; j < n; j++) {
      const double lj = a[j * lda + j];
      rank += (lj < li) || (lj == li && j < i);
    }
    w[(long long)b * n + rank] = li;
    for (int r = 0; r < n; r++) V[(long long)b * n * n + (long long)r * n + rank] = vt[(long long)i * n + r];
  }
}

// ---------------------------------------------------------------------------------------------
// d(sum gK o K)/d(inv_ls, amp), K rectangular [n1][n2]; partial per 64x64 tile, then grad_finish (dense_kernels.cuh)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kernel_bwd_kernel(const double* __restrict__ x1, const double* __restrict__ x2,
                                                         const double* __restrict__ w, const double* __restrict__ amp,
                                                         const double* __restrict__ gK, int n1, int n2, int d,
                                                         double* __restrict__ partial) {
  extern __shared__ __align__(16) double gsm[];
  const int ldx = d + 1;
  double* xi = gsm;
  double* xj = xi + 64 * ldx;
  double* red = xj + 64 * ldx;      // [8][d+1]
  const int tid = threadIdx.x, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  for (int e = tid; e < 2 * 64 * d; e += 256) {
    const int which = e / (64 * d), r = (e / d) % 64, k = e % d;
    const int gi = (which ? j0 : i0) + r, nn = which ? n2 : n1;
    const double* xx = which ? x2 : x1;
    (which ? xj : xi)[r * ldx + k] = (gi < nn) ? xx[(long long)gi * d + k] * w[k] : 0.0;
  }
  __syncthreads();
  const int tr = tid >> 4, tc = tid & 15;
  const double a = amp[0];
  double Wv[4][4], sumW = 0.0;
#pragma unroll
  for (int u = 0; u < 4; u++)
#pragma unroll
    for (int v = 0; v < 4; v++) {
      const int r = tr + 16 * u, c = tc + 16 * v, gi = i0 + r, gj = j0 + c;
      double wgt = 0.0;
      if (gi < n1 && gj < n2) {
        double sq = 0.0;
        for (int k = 0; k < d; k++) { const double dz = xi[r * ldx + k] - xj[c * ldx + k]; sq = fma(dz, dz, sq); }
        wgt = gK[(long long)gi * n2 + gj] * exp(-0.5 * sq);
      }
      sumW += wgt;
      Wv[u][v] = wgt * a;
    }
  const int warp = tid >> 5, lane = tid & 31;
  for (int k = 0; k < d; k++) {
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int v = 0; v < 4; v++) {
        const double dz = xi[(tr + 16 * u) * ldx + k] - xj[(tc + 16 * v) * ldx + k];
        acc = fma(Wv[u][v] * dz, dz, acc);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp * (d + 1) + k] = acc;
  }
  {
    double v = sumW;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * (d + 1) + d] = v;
  }
  __syncthreads();
  double* part = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * (d + 1);
  for (int k = tid; k <= d; k += 256) {
    double s = 0.0;
    for (int wv = 0; wv < 8; wv++) s += red[wv * (d + 1) + k];
    part[k] = s;
  }
}

__global__ void kernel_bwd_finish_kernel(const double* __restrict__ partial, int npart, int d, const double* __restrict__ w,
                                         const double* __restrict__ amp, double* __restrict__ g_w, double* __restrict__ g_amp) {
  for (int k = threadIdx.x; k <= d; k += blockDim.x) {
    double s = 0.0;
    for (int t = 0; t < npart; t++) s += partial[(long long)t * (d + 1) + k];
    if (k < d) g_w[k] = -s / w[k];
    else g_amp[0] = s;
  }
}

static KronSizes make_sizes(const int* sizes, int nmodes) {
  KronSizes s;
  memset(&s, 0, sizeof(s));
  s.nmodes = nmodes;
  int off = 0;
  for (int m = 0; m < nmodes; m++) { s.n[m] = sizes[m]; s.off[m] = off; off += sizes[m]; }
  return s;
}

}  // namespace ffgp

using namespace ffgp;

extern "C" {

int ffgp_mode_dot_f64(const double* t, const double* mat, double* out, long long outer, int I, long long inner, int J,
                      int transpose_mat, void* stream) {
  if (!t || !mat || !out) return fail(-1, "ffgp_mode_dot_f64: null pointer%s", "");
  if (outer <= 0 || I <= 0 || inner <= 0 || J <= 0) return fail(-2, "ffgp_mode_dot_f64: bad size%s", "");
  const long long ncols = outer * inner;
  const long long nchunks = (ncols + MD_COLS - 1) / MD_COLS;
  const int gx = (int)std::min<long long>(nchunks, 148LL * 8);
  mode_dot_kernel<<<dim3(gx, (J + 127) / 128), 256, 0, (cudaStream_t)stream>>>(t, mat, out, ncols, I, inner, J, transpose_mat);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  return 0;
}

static int gram_slices(long long ncols, int Ja, int Jb) {
  const long long tiles = (long long)((Ja + MG_T - 1) / MG_T) * ((Jb + MG_T - 1) / MG_T);
  long long want = std::max<long long>(1, (148LL * 4) / tiles);
  long long maxs = std::max<long long>(1, ncols / 256);
  return (int)std::min<long long>(std::min(want, maxs), 1024);
}

size_t ffgp_mode_gram_scratch_bytes(long long outer, long long inner, int Ja, int Jb) {
  return (size_t)gram_slices(outer * inner, Ja, Jb) * Ja * Jb * sizeof(double) + 256;
}

int ffgp_mode_gram_f64(const double* X, const double* Y, double* G, long long outer, long long inner, int Ja, int Jb,
                       void* scratch, size_t scratch_bytes, void* stream) {
  if (!X || !Y || !G || !scratch) return fail(-1, "ffgp_mode_gram_f64: null pointer%s", "");
  if (outer <= 0 || inner <= 0 || Ja <= 0 || Jb <= 0) return fail(-2, "ffgp_mode_gram_f64: bad size%s", "");
  if (scratch_bytes < ffgp_mode_gram_scratch_bytes(outer, inner, Ja, Jb)) return fail(-3, "ffgp_mode_gram_f64: scratch too small%s", "");
  const long long ncols = outer * inner;
  const int ns = gram_slices(ncols, Ja, Jb);
  long long cps = (ncols + ns - 1) / ns;
  cps = (cps + MG_KC - 1) / MG_KC * MG_KC;
  cudaStream_t st = (cudaStream_t)stream;
  mode_gram_kernel<<<dim3((Jb + MG_T - 1) / MG_T, (Ja + MG_T - 1) / MG_T, ns), 256, 0, st>>>(X, Y, (double*)scratch, ncols, inner,
                                                                                             Ja, Jb, cps);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  const long long nelem = (long long)Ja * Jb;
  gram_reduce_kernel<<<(unsigned)((nelem + 255) / 256), 256, 0, st>>>((const double*)scratch, ns, nelem, G);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  return 0;
}

size_t ffgp_kron_core_scratch_bytes(long long total) {
  (void)total;
  return (size_t)148 * 8 * 4 * sizeof(double) + 256;
}

int ffgp_kron_core_f64(const double* T1, const double* lambdas, const int* sizes_host, int nmodes, const double* noise_inv,
                       double add_scalar, double* out_core, double* out_A, double* out_sums, void* scratch,
                       size_t scratch_bytes, void* stream) {
  if (!T1 || !lambdas || !sizes_host || !out_sums || !scratch) return fail(-1, "ffgp_kron_core_f64: null pointer%s", "");
  if (nmodes <= 0 || nmodes > 8) return fail(-2, "ffgp_kron_core_f64: 1..8 modes supported%s", "");
  if (scratch_bytes < ffgp_kron_core_scratch_bytes(0)) return fail(-3, "ffgp_kron_core_f64: scratch too small%s", "");
  KronSizes s = make_sizes(sizes_host, nmodes);
  long long total = 1;
  for (int m = 0; m < nmodes; m++) total *= sizes_host[m];
  const int nb = (int)std::min<long long>((total + 255) / 256, 148LL * 8);
  cudaStream_t st = (cudaStream_t)stream;
  kron_core_kernel<<<nb, 256, 0, st>>>(T1, lambdas, s, noise_inv, add_scalar, total, out_core, out_A, (double*)scratch);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  kron_sums_finish_kernel<<<1, 32, 0, st>>>((const double*)scratch, nb, out_sums);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  return 0;
}

int ffgp_kron_scale_f64(const double* in, const double* lambdas, const int* sizes_host, int nmodes, int skip_mode,
                        int divide_by_A, const double* noise_inv, double add_scalar, double* out, void* stream) {
  if (!lambdas || !sizes_host || !out) return fail(-1, "ffgp_kron_scale_f64: null pointer%s", "");
  if (nmodes <= 0 || nmodes > 8) return fail(-2, "ffgp_kron_scale_f64: 1..8 modes supported%s", "");
  KronSizes s = make_sizes(sizes_host, nmodes);
  long long total = 1;
  for (int m = 0; m < nmodes; m++) total *= sizes_host[m];
  const int nb = (int)std::min<long long>((total + 255) / 256, 148LL * 8);
  kron_scale_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(in, lambdas, s, skip_mode, divide_by_A, noise_inv, add_scalar, total, out);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  return 0;
}

size_t ffgp_syevj_workspace_bytes(int n, int batch) {
  if (n <= 0 || batch <= 0) return 0;
  return (size_t)batch * n * n * sizeof(double) * (n <= SYEVJ_SMEM_N ? 1 : 2) + 256;
}

int ffgp_syevj_f64(const double* A, int n, int batch, double* w, double* V, void* workspace, size_t workspace_bytes,
                   int* info, void* stream) {
  if (!A || !w || !V || !workspace || !info) return fail(-1, "ffgp_syevj_f64: null pointer%s", "");
  if (n <= 0 || n > 512 || batch <= 0) return fail(-2, "ffgp_syevj_f64: n must be in 1..512%s", "");
  if (workspace_bytes < ffgp_syevj_workspace_bytes(n, batch)) return fail(-3, "ffgp_syevj_f64: workspace too small%s", "");
  cudaStream_t st = (cudaStream_t)stream;
  const int m = (n + 1) & ~1;
  const bool in_smem = n <= SYEVJ_SMEM_N;
  size_t smem = (size_t)(m + SYEVJ_THREADS) * sizeof(double);
  if (in_smem) smem += (size_t)n * (n + 1) * sizeof(double);
  static bool attr = false;
  if (!attr) {
    FFGP_CUDA(cudaFuncSetAttribute(syevj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  double* vt = (double*)workspace;
  double* wa = in_smem ? nullptr : vt + (size_t)batch * n * n;
  FFGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * batch, st));
  syevj_kernel<<<batch, SYEVJ_THREADS, smem, st>>>(A, n, w, V, wa, vt, info, 40);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  return 0;
}

size_t ffgp_kernel_matrix_bwd_scratch_bytes(int n1, int n2, int d, int batch) {
  if (n1 <= 0 || n2 <= 0 || d <= 0 || batch <= 0) return 0;
  return (size_t)((n1 + 63) / 64) * ((n2 + 63) / 64) * (d + 1) * sizeof(double) * batch + 256;
}

int ffgp_kernel_matrix_bwd_f64(const double* x1, const double* x2, const double* inv_ls, const double* amp, const double* gK,
                               int n1, int n2, int d, int batch, int params_batched, double* g_inv_ls, double* g_amp,
                               void* scratch, size_t scratch_bytes, void* stream) {
  if (!x1 || !x2 || !inv_ls || !amp || !gK || !g_inv_ls || !g_amp || !scratch)
    return fail(-1, "ffgp_kernel_matrix_bwd_f64: null pointer%s", "");
  if (n1 <= 0 || n2 <= 0 || d <= 0 || d > 64 || batch <= 0) return fail(-2, "ffgp_kernel_matrix_bwd_f64: bad size (d <= 64)%s", "");
  if (scratch_bytes < ffgp_kernel_matrix_bwd_scratch_bytes(n1, n2, d, batch)) return fail(-3, "ffgp_kernel_matrix_bwd_f64: scratch too small%s", "");
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    FFGP_CUDA(cudaFuncSetAttribute(kernel_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  const int gx = (n2 + 63) / 64, gy = (n1 + 63) / 64;
  const size_t smem = (size_t)(2 * 64 * (d + 1) + 8 * (d + 1)) * sizeof(double);
  for (int b = 0; b < batch; b++) {
    double* part = (double*)scratch + (size_t)b * gx * gy * (d + 1);
    const double* wv = inv_ls + (params_batched ? (size_t)b * d : 0);
    const double* av = amp + (params_batched ? b : 0);
    kernel_bwd_kernel<<<dim3(gx, gy), 256, smem, st>>>(x1 + (size_t)b * n1 * d, x2 + (size_t)b * n2 * d, wv, av,
                                                       gK + (size_t)b * n1 * n2, n1, n2, d, part);
    ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
    kernel_bwd_finish_kernel<<<1, 128, 0, st>>>(part, gx * gy, d, wv, av, g_inv_ls + (size_t)b * d, g_amp + b);
    ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
