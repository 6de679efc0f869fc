// Device kernels of the dense GP path other than the level-3 GEMM (gemm_dmma.cuh):
//   kernel_matrix_kernel    fused ||x||^2+||y||^2-2xy^T (DMMA) + ARD scaling + exp + diag/noise epilogue
//   potrf_trtri_base_kernel 128x128 diagonal block: Cholesky factor AND its inverse in one sweep
//   trmv_lower_kernel       Gamma = M Y            (M = L^-1 lower, few right-hand sides, HBM-bound)
//   colsum_weighted_kernel  out = Mat^T W          (alpha = M^T Gamma, mean = K*^T alpha, colsum V^2)
//   nll_reduce_kernel       0.5 ||Gamma||^2 + D sum log L_ii, fixed-order reduction
//   grad_contract_kernel    G o K contractions -> d/d inv_ls, d/d amp, diag(G)
//   pad_copy_kernel         padded <-> user layouts
// All matrices are row-major; "np" is n rounded up to a multiple of 128 (identity tail).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>
#include "gemm_dmma.cuh"

namespace ffgp {

// ------------------------------------------------------------------------------------------
// Kernel-matrix assembly.  K[i][j] = amp * exp(-0.5 * sum_k ((x1[i][k]-x2[j][k]) * w[k])^2) + offset
//   (+ diag_add[i] on the diagonal, + sigma_add[i][j]) on the real n1 x n2 part; the padded tail is
//   the identity (symmetric case) or zero.  The cross term runs on the FP64 tensor pipe.
// ------------------------------------------------------------------------------------------
struct KernelMatrixParams {
  const double* x1; const double* x2;     // [batch][n1][d], [batch][n2][d]
  const double* w;                        // inverse length scales [d] (+ batch stride sw)
  const double* amp;                      // [1] (+ batch stride samp); NULL => K = 0 (covariance given in sigma_add)
  const double* diag_add;                 // [n1] or NULL (+ batch stride sdiag)
  const double* offset;                   // [1] or NULL  (+ batch stride soff): added to every real entry
  const double* sigma_add;                // [n1][n2] or NULL (+ batch stride ssig)
  double* K;                              // [batch][np1][ldk]
  int n1, n2, d, np1, np2, ldk;
  long long sx1, sx2, sw, samp, sdiag, soff, ssig, sK;
  int symmetric;                          // x1 == x2: identity tail
  int lower_only;                         // symmetric: skip tiles strictly above the diagonal
  int clamp;                              // clamp the squared distance at 0 (torch.cdist semantics)
  int bounded;                            // K is the caller's exact [n1][n2] array: predicated scalar stores
};

// exp(x) for x <= 0 (and small positive x), branch-free: n = rint(x log2 e) by the 1.5 * 2^52 trick, r = x - n ln 2 in two pieces (|r| <= 0.347),
// degree-13 Taylor polynomial (truncation (ln 2 / 2)^14 / 14! = 4e-18), scaling through the exponent field; results
// below 2^-1020 are flushed to zero.  Error <= 1 ulp of the result like exp(), without its special-case paths, so the
// 16 evaluations of a thread stay straight-line code.
__device__ __forceinline__ double exp_nonpos(double x) {
  const double MAGIC = 6755399441055744.0;
  x = fmax(x, -800.0);                               // exp(-800) is flushed to zero below; keeps n in int range
  const double t = fma(x, 1.4426950408889634074, MAGIC);
  const int n = __double2loint(t);
  const double nf = t - MAGIC;
  double r = fma(nf, -6.93147180369123816490e-01, x);
  r = fma(nf, -1.90821492927058770002e-10, r);
  double q = 1.6059043836821614599e-10;              // 1/13!
  q = fma(q, r, 2.0876756987868098979e-09);          // 1/12!
  q = fma(q, r, 2.5052108385441718775e-08);          // 1/11!
  q = fma(q, r, 2.7557319223985890653e-07);          // 1/10!
  q = fma(q, r, 2.7557319223985892511e-06);          // 1/9!
  q = fma(q, r, 2.4801587301587301566e-05);          // 1/8!
  q = fma(q, r, 1.9841269841269841253e-04);          // 1/7!
  q = fma(q, r, 1.3888888888888889419e-03);          // 1/6!
  q = fma(q, r, 8.3333333333333332177e-03);          // 1/5!
  q = fma(q, r, 4.1666666666666664354e-02);          // 1/4!
  q = fma(q, r, 1.6666666666666665741e-01);          // 1/3!
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  const double scale = __hiloint2double((max(n, -1021) + 1023) << 20, 0);
  return n < -1020 ? 0.0 : q * scale;
}

__global__ void __launch_bounds__(256) kernel_matrix_kernel(const KernelMatrixParams p) {
  constexpr int T = 64, DC = 16, LD = DC + 4;
  const int ti = blockIdx.y, tj = blockIdx.x, b = blockIdx.z;
  if (p.lower_only && tj > ti) return;
  __shared__ double xs1[T][LD], xs2[T][LD];
  __shared__ double nrm[2][T];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = warp & 3;        // 2 x 4 warps, warp tile 32 x 16
  const int g = lane >> 2, tq = lane & 3;
  const double* x1 = p.x1 ? p.x1 + b * p.sx1 : nullptr;
  const double* x2 = p.x2 ? p.x2 + b * p.sx2 : nullptr;
  const double* w = p.w ? p.w + b * p.sw : nullptr;
  const int i0 = ti * T, j0 = tj * T;
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
  if (tid < 2 * T) nrm[tid / T][tid % T] = 0.0;
  const bool have_k = (p.amp != nullptr);
  if (have_k) {
    for (int k0 = 0; k0 < p.d; k0 += DC) {
      __syncthreads();
      for (int e = tid; e < 2 * T * DC; e += 256) {
        const int which = e / (T * DC), r = (e / DC) % T, k = e % DC;
        const int gi = (which ? j0 : i0) + r, gk = k0 + k;
        const int nn = which ? p.n2 : p.n1;
        const double* xx = which ? x2 : x1;
        double v = 0.0;
        if (gi < nn && gk < p.d) v = xx[(long long)gi * p.d + gk] * w[gk];
        (which ? xs2 : xs1)[r][k] = v;
      }
      __syncthreads();
      if (tid < 2 * T) {
        const int which = tid / T, r = tid % T;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DC; k++) { const double v = (which ? xs2 : xs1)[r][k]; s = fma(v, v, s); }
        nrm[which][r] += s;
      }
#pragma unroll
      for (int kk = 0; kk < DC / 4; kk++) {
        double af[4], bf[2];
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = xs1[wm * 32 + i * 8 + g][kk * 4 + tq];
#pragma unroll
        for (int j = 0; j < 2; j++) bf[j] = xs2[wn * 16 + j * 8 + g][kk * 4 + tq];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
  }
  __syncthreads();
  const double amp = have_k ? p.amp[b * p.samp] : 0.0;
  const double off = p.offset ? p.offset[b * p.soff] : 0.0;
  const double* dg = p.diag_add ? p.diag_add + b * p.sdiag : nullptr;
  const double* sg = p.sigma_add ? p.sigma_add + b * p.ssig : nullptr;
  double* K = p.K + b * p.sK;
  // all 16 kernel values of the thread first, straight-line (the exp chains interleave), then the optional terms
  double kv[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        double sq = nrm[0][wm * 32 + i * 8 + g] + nrm[1][wn * 16 + j * 8 + tq * 2 + e] - 2.0 * acc[i][j][e];
        if (p.clamp) sq = fmax(sq, 0.0);
        kv[i][j][e] = have_k ? fma(amp, exp_nonpos(-0.5 * sq), off) : off;
      }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = wm * 32 + i * 8 + g, gi = i0 + r;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int c = wn * 16 + j * 8 + tq * 2;
      double out[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int gj = j0 + c + e;
        double v = kv[i][j][e];
        if (gi < p.n1 && gj < p.n2) {
          if (sg) v += sg[(long long)gi * p.n2 + gj];
          if (dg && gi == gj) v += dg[gi];
        } else {
          v = (p.symmetric && gi == gj) ? 1.0 : 0.0;
        }
        out[e] = v;
      }
      if (p.bounded) {
        if (gi < p.n1) {
          if (j0 + c < p.n2) K[(long long)gi * p.ldk + j0 + c] = out[0];
          if (j0 + c + 1 < p.n2) K[(long long)gi * p.ldk + j0 + c + 1] = out[1];
        }
      } else {
        *reinterpret_cast<double2*>(K + (long long)gi * p.ldk + j0 + c) = make_double2(out[0], out[1]);
      }
    }
  }
}

// Assembly of the symmetric kernel matrix of a factorisation (x1 == x2, lower part, padded to a multiple of 128 with an
// identity tail, d <= 16, no sigma_add / offset): one CTA per LOWER 128 x 128 super-tile, its four 64 x 64 sub-tiles in a
// rolled loop over the same code.  Why a second kernel: `ncu --set full` of kernel_matrix_kernel on BASELINE config 5
// (gpurun_out/r02_c5_small_kernels.ncu-rep) shows it ISSUE-bound - 66 % issue-slot utilisation at 35 % FP64-pipe
// utilisation, ~105 instructions per matrix entry of which ~30 are FP64 (the cross term on DMMA, one exp): the rest is
// the per-entry bounds / diagonal / optional-term logic of the general kernel, the staging of the scaled inputs (redone
// for every 64 x 64 tile) and 44 % of the grid being CTAs above the diagonal that exit at once.  Here the inputs of 128
// rows and 128 columns are staged once per 16384 entries, interior super-tiles skip every per-entry check, and the grid
// enumerates lower super-tiles only.
__global__ void __launch_bounds__(256) kernel_matrix_sym128_kernel(const KernelMatrixParams p) {
  constexpr int T = 128, S = 64, DC = 16, LD = DC + 4;
  int t = blockIdx.x;
  int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (ti * (ti + 1) / 2 > t) --ti;
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  const int tj = t - ti * (ti + 1) / 2, b = blockIdx.y;
  __shared__ double xs1[T][LD], xs2[T][LD];
  __shared__ double nrm[2][T];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = warp & 3;        // 2 x 4 warps, warp tile 32 x 16 of a 64 x 64 sub-tile
  const int g = lane >> 2, tq = lane & 3;
  const double* x = p.x1 + b * p.sx1;
  const double* w = p.w + b * p.sw;
  const int i0 = ti * T, j0 = tj * T, n = p.n1, d = p.d;
  for (int e = tid; e < 2 * T * DC; e += 256) {
    const int which = e / (T * DC), r = (e / DC) % T, k = e % DC;
    const int gi = (which ? j0 : i0) + r;
    double v = 0.0;
    if (gi < n && k < d) v = x[(long long)gi * d + k] * w[k];
    (which ? xs2 : xs1)[r][k] = v;
  }
  __syncthreads();
  {
    const int which = tid / T, r = tid % T;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < DC; k++) { const double v = (which ? xs2 : xs1)[r][k]; s = fma(v, v, s); }
    nrm[which][r] = s;
  }
  __syncthreads();
  const double amp = p.amp[b * p.samp];
  const double* dg = p.diag_add ? p.diag_add + b * p.sdiag : nullptr;
  double* K = p.K + b * p.sK;
  const bool interior = (i0 + T <= n) && (j0 + T <= n);
  const int nkk = (d + 3) >> 2;                    // k4 steps that hold non-zero dimensions
#pragma unroll 1
  for (int st = 0; st < 4; st++) {
    const int si = st >> 1, sj = st & 1;
    if (ti == tj && sj > si) continue;             // above the diagonal
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
#pragma unroll 1
    for (int kk = 0; kk < nkk; kk++) {
      double af[4], bf[2];
#pragma unroll
      for (int i = 0; i < 4; i++) af[i] = xs1[si * S + wm * 32 + i * 8 + g][kk * 4 + tq];
#pragma unroll
      for (int j = 0; j < 2; j++) bf[j] = xs2[sj * S + wn * 16 + j * 8 + g][kk * 4 + tq];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    double kv[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          double sq = nrm[0][si * S + wm * 32 + i * 8 + g] + nrm[1][sj * S + wn * 16 + j * 8 + tq * 2 + e] - 2.0 * acc[i][j][e];
          if (p.clamp) sq = fmax(sq, 0.0);
          kv[i][j][e] = amp * exp_nonpos(-0.5 * sq);
        }
    const bool diag_sub = (ti == tj) && (si == sj);
    if (interior && !diag_sub) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int gi = i0 + si * S + wm * 32 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int gj = j0 + sj * S + wn * 16 + j * 8 + tq * 2;
          *reinterpret_cast<double2*>(K + (long long)gi * p.ldk + gj) = make_double2(kv[i][j][0], kv[i][j][1]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int gi = i0 + si * S + wm * 32 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int gj0 = j0 + sj * S + wn * 16 + j * 8 + tq * 2;
          double out[2];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int gj = gj0 + e;
            double v = kv[i][j][e];
            if (gi < n && gj < n) { if (dg && gi == gj) v += dg[gi]; }
            else v = (gi == gj) ? 1.0 : 0.0;
            out[e] = v;
          }
          *reinterpret_cast<double2*>(K + (long long)gi * p.ldk + gj0) = make_double2(out[0], out[1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// 128x128 diagonal block: L = chol(A_block) and M = L^-1 in ONE right-looking sweep.
// The forward substitution L M = I is carried along with the factorisation (the rank-8 update
// of panel p is applied to the rows below it in BOTH the trailing part of A and the running
// inverse), so potrf and trtri share every pass over shared memory.
// smem T[128][129]: T[i][k], k<=i holds A/L; T[c][i+1], c<=i holds R[i][c] (running inverse).
// The odd row stride makes both row-wise and column-wise warp accesses bank-conflict free.
//
// This kernel sits on the critical path of the blocked Cholesky (one CTA, everything else waits for it), and what
// bounds it is the dependent chain  rsqrt -> scale -> update -> rsqrt  of the 128 pivots (about 150 clk per column),
// so the work is split by role (profiles/r01_base_kernel_timeline.txt: the previous version, where every warp did a
// share of everything behind divergent branches, needed 8400 clk per 8-column panel; 2000 is the chain):
//   warp 0      the critical chain of panel p+1 only: panel solve of the 8 rows of the next diagonal block, their
//               rank-8 update of that 8x8 block, its Cholesky in registers (every lane redundantly - no exchange on
//               the chain), publish the factor for the next iteration;
//   warps 1-7   everything off the chain for panel p: panel solve of the remaining rows and of the inverse's row
//               block (one thread per row / column), then the rank-8 update of the trailing matrix and of the running
//               inverse (a warp per 4-row group, lane owns 4 "virtual columns": v < r0 inverse, v >= r0 matrix).
// One CTA barrier per panel plus one producer/consumer named barrier (warp 0 arrives after its panel rows are stored,
// warps 1-7 wait before the update).  All loads are unconditional with selected addresses and all masked stores go
// to a dummy word: no divergent branches in the loop.
// ------------------------------------------------------------------------------------------
constexpr int BASE_N = 128;                                  // block size of the single-problem path (and of the padding)
constexpr int BASE_N_BATCHED = 64;                           // experimental batched base block (see factor_rec); also the log-det slot size
constexpr int BASE_F = 48;                                   // 36 factor entries + 8 inverses + fail column (+pad)
constexpr size_t base_smem_bytes(int bn) { return ((size_t)bn * (bn + 1) + 2 * BASE_F + 64 + 64 + 8 + 256) * sizeof(double); }
constexpr size_t BASE_SMEM = base_smem_bytes(BASE_N);

#ifdef FFGP_BASE_TRACE
__device__ long long g_base_trace[8 * 16 * 4];          // [warp][panel][slot] clock64 stamps (tools/base_trace.cu)
#define BASE_STAMP(panel, slot) do { if (lane == 0) g_base_trace[(warp * 16 + (panel)) * 4 + (slot)] = clock64(); } while (0)
#else
#define BASE_STAMP(panel, slot) do { } while (0)
#endif

// Cholesky of an 8x8 block held in registers (lower part of l), returns first failing column or -1.
__device__ __forceinline__ int chol8_regs(double (&l)[8][8], double (&inv)[8]) {
  // Right-looking (outer-product) order: as soon as column c is scaled, its rank-1 update is applied to the columns to
  // its right.  Every entry still receives its updates in the order k = 0, 1, .. (bitwise the same result as the
  // dot-product form), but the dependent chain per column shrinks from (2 c + 13) to ~14 FP64 operations - this routine
  // IS the critical path of a late panel (profiles/r01_base_kernel_timeline_v2.txt: 2100 of ~2900 clk).
  int fail = -1;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const double s = l[c][c];
    if (!(s > 0.0) && fail < 0) fail = c;
    // one rsqrt replaces sqrt + division on the serial chain (l_cc = s * rsqrt(s), 1/l_cc = rsqrt(s))
    const double rs = rsqrt(s);
    inv[c] = rs;
    l[c][c] = s * rs;
#pragma unroll
    for (int r = c + 1; r < 8; r++) l[r][c] *= rs;
#pragma unroll
    for (int r = c + 1; r < 8; r++)
#pragma unroll
      for (int k = c + 1; k <= r; k++) l[r][k] = fma(-l[r][c], l[k][c], l[r][k]);
  }
  return fail;
}

// a <- a L^-T for one row a[0..7] against the 8x8 factor (l, inv): right-looking order, same rounding as the
// dot-product form, dependent chain 2 operations per column instead of c + 1
__device__ __forceinline__ void solve8_row(double (&a)[8], const double (&l)[8][8], const double (&inv)[8]) {
#pragma unroll
  for (int c = 0; c < 8; c++) {
    a[c] *= inv[c];
#pragma unroll
    for (int k = c + 1; k < 8; k++) a[k] = fma(-a[c], l[k][c], a[k]);
  }
}

__device__ __forceinline__ void base_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void base_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// The factorisation proper, on a block that is already in shared memory: sm = T[BN][BN + 1] (lower part A, the inverse
// part initialised to the identity) followed by the scratch area laid out below (base_smem_bytes).  All 256 threads of
// the CTA call it; T is valid for every thread after the caller's next __syncthreads().  Returns the first failing
// column (every thread tracks the same value) or -1.  Shared by potrf_trtri_base_kernel and factor256_kernel.
template <int BN>
__device__ __forceinline__ int base_factor_smem(double* __restrict__ sm, const int tid) {
  constexpr int BASE_N = BN, BASE_LD = BN + 1, QV = BN / 32;       // QV: 32-wide groups of virtual columns per lane
  static_assert(BN == 64 || BN == 128, "base block is 64 or 128");
  double* T = sm;                                   // [BN][BN + 1]
  double* F = sm + BASE_N * BASE_LD;                // [2][BASE_F] published diagonal factors (double-buffered)
  double* Dn = F + 2 * BASE_F;                      // [36] updated next diagonal block (warp 0 scratch)
  // sink for masked stores / source of masked loads: one slot PER THREAD, so that the branch-free masking below is not a
  // (benign) shared-memory race between threads - compute-sanitizer racecheck stays clean (profiles/r02_sanitizer.txt)
  const int warp = tid >> 5, lane = tid & 31;
  const int DUMMY = BASE_N * BASE_LD + 2 * BASE_F + 136 + tid;
#define TT(i, k) T[(i) * BASE_LD + (k)]
  int fail_col = -1;
  double l[8][8], inv[8];
  {  // factor of the first diagonal block: every thread redundantly, thread 0 stores it
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int c = 0; c <= r; c++) l[r][c] = TT(r, c);
    const int f = chol8_regs(l, inv);
    if (f >= 0) fail_col = f;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) TT(r, c) = l[r][c];
    }
  }
  // warp 0: (row, col) of the next-diagonal-block entries a lane updates (entry e and, for the tail, entry 32 + e % 4)
  int ea0, eb0, ea1, eb1;
  {
    auto decode = [](int e, int& a, int& bb) {
      a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
      if (a * (a + 1) / 2 > e) --a;
      if ((a + 1) * (a + 2) / 2 <= e) ++a;
      bb = e - a * (a + 1) / 2;
    };
    decode(lane, ea0, eb0);
    decode(32 + (lane & 3), ea1, eb1);
  }
  for (int p = 0; p < BASE_N / 8; p++) {
    const int j0 = p * 8, r0 = j0 + 8;
    if (p > 0) {
      __syncthreads();
    }
    if (p > 0 && warp != 0) {                       // warp 0 carries the factor it computed in registers
      const double* Fp = F + (p & 1) * BASE_F;
      int q = 0;
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) l[r][c] = Fp[q++];
#pragma unroll
      for (int c = 0; c < 8; c++) inv[c] = Fp[36 + c];
      const int f = (int)Fp[44];
      if (f >= 0 && fail_col < 0) fail_col = j0 + f;
    }
    BASE_STAMP(p, 0);
    if (warp == 0) {
      if (r0 == BASE_N) continue;                  // last panel: nothing below it
      // ---- W0.1: panel solve of the rows of the NEXT diagonal block:  a <- a L11^-T ----
      if (lane < 8) {
        const int r = r0 + lane;
        double a[8];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = TT(r, j0 + c);
        solve8_row(a, l, inv);
#pragma unroll
        for (int c = 0; c < 8; c++) TT(r, j0 + c) = a[c];
      }
      __syncwarp();
      __threadfence_block();
      base_bar_arrive(1, 256);                     // warps 1-7 may now read these 8 panel rows
      BASE_STAMP(p, 1);
      // ---- W0.2: rank-8 update of the next 8x8 diagonal block: lane e owns entry e, lanes 0-3 also entry 32+e
      //      (the other lanes recompute entry e there, harmlessly, so the two chains interleave without divergence) ----
      {
        double d0 = TT(r0 + ea0, r0 + eb0), d1 = TT(r0 + ea1, r0 + eb1);
        const double* pa0 = T + (r0 + ea0) * BASE_LD + j0;
        const double* pb0 = T + (r0 + eb0) * BASE_LD + j0;
        const double* pa1 = T + (r0 + ea1) * BASE_LD + j0;
        const double* pb1 = T + (r0 + eb1) * BASE_LD + j0;
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          d0 = fma(-pa0[kk], pb0[kk], d0);
          d1 = fma(-pa1[kk], pb1[kk], d1);
        }
        Dn[lane] = d0;
        Dn[32 + (lane & 3)] = d1;                  // lanes sharing (lane & 3) write identical values
      }
      __syncwarp();
      BASE_STAMP(p, 2);
      // ---- W0.3: its Cholesky, every lane redundantly in registers; publish for the next iteration ----
      {
        int q = 0;
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int c = 0; c <= r; c++) l[r][c] = Dn[q++];
      }
      const int f = chol8_regs(l, inv);
      if (f >= 0 && fail_col < 0) fail_col = r0 + f;
      double* Fn = F + ((p + 1) & 1) * BASE_F;
      if (lane == 0) {
        int q = 0;
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int c = 0; c <= r; c++) Fn[q++] = l[r][c];
#pragma unroll
        for (int c = 0; c < 8; c++) Fn[36 + c] = inv[c];
        Fn[44] = (double)f;
      } else if (lane == 1) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int c = 0; c <= r; c++) TT(r0 + r, r0 + c) = l[r][c];
      }
      BASE_STAMP(p, 3);
      continue;
    }
    // =============================== warps 1-7 ===============================
    // ---- O.1: panel solve, one item per thread: rows below the next diagonal block, then the inverse's columns ----
    {
      const int item = tid - 32;
      const int nrows = BASE_N - r0 - 8;           // rows r0+8 .. 127  (may be <= 0)
      if (item < nrows) {
        const int r = r0 + 8 + item;
        double a[8];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = TT(r, j0 + c);
        solve8_row(a, l, inv);
#pragma unroll
        for (int c = 0; c < 8; c++) TT(r, j0 + c) = a[c];
      } else {
        const int c = item - max(nrows, 0);
        if (c < j0 + 8) {                          // inverse row block: v <- L11^-1 v   (column c of rows j0..j0+7)
          double v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const double x = T[(c <= j0 + k) ? c * BASE_LD + j0 + k + 1 : DUMMY];
            v[k] = (c <= j0 + k) ? x : 0.0;
          }
#pragma unroll
          for (int k = 0; k < 8; k++) {                 // right-looking forward substitution (same rounding, short chain)
            v[k] *= inv[k];
#pragma unroll
            for (int q = k + 1; q < 8; q++) v[q] = fma(-l[q][k], v[k], v[q]);
          }
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const int idx = (c <= j0 + k) ? c * BASE_LD + j0 + k + 1 : DUMMY;
            T[idx] = v[k];
          }
        }
      }
    }
    if (r0 == BASE_N) continue;
    base_bar_sync(1, 256);
    BASE_STAMP(p, 1);
    // ---- O.2: rank-8 update of the rows below the panel.  Lane owns virtual columns v = lane + 32 q. ----
    const int nquad = (BASE_N - r0) >> 2;
    double bv[8][QV];
#pragma unroll
    for (int q = 0; q < QV; q++) {
      const int v = lane + 32 * q;
      const int isR = (v < r0) ? 1 : 0;            // running inverse (row block of the panel) vs panel rows of L
      const double* src = T + v * BASE_LD + j0 + isR;
#pragma unroll
      for (int kk = 0; kk < 8; kk++) {
        const double x = src[kk];
        bv[kk][q] = (isR && v > j0 + kk) ? 0.0 : x;
      }
    }
    BASE_STAMP(p, 2);
    // A row group only has entries in the 32-wide groups of virtual columns that start at or below its last row
    // (v <= i): the groups beyond are all masked lanes - at panel 0 more than half of the FMAs of the original
    // all-groups loop (profiles/r02_base_kernel_timeline_v3.txt).  One body per group count, selected per row group.
    auto quad = [&](auto qn_tag, const int rq) {
      constexpr int QN = decltype(qn_tag)::value;
      const int i0 = r0 + 4 * rq;
      double cacc[4][QN];
      int idx[4][QN];
#pragma unroll
      for (int q = 0; q < QN; q++) {
        const int v = lane + 32 * q;
        const bool isR = v < r0;
#pragma unroll
        for (int a = 0; a < 4; a++) {
          const int i = i0 + a;
          const int id = isR ? v * BASE_LD + i + 1 : i * BASE_LD + v;
          // the 8x8 block on the diagonal right after the panel belongs to warp 0; entries above the diagonal do not exist:
          // both the load and the store of a masked entry go to this thread's private slot
          idx[a][q] = (isR || (v <= i && i >= r0 + 8)) ? id : DUMMY;
          cacc[a][q] = T[idx[a][q]];
        }
      }
      double av[4][8];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int kk = 0; kk < 8; kk++) av[a][kk] = -TT(i0 + a, j0 + kk);
#pragma unroll
      for (int kk = 0; kk < 8; kk++) {
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int q = 0; q < QN; q++) cacc[a][q] = fma(av[a][kk], bv[kk][q], cacc[a][q]);
      }
#pragma unroll
      for (int q = 0; q < QN; q++)
#pragma unroll
        for (int a = 0; a < 4; a++) T[idx[a][q]] = cacc[a][q];
    };
    for (int rq = warp - 1; rq < nquad; rq += 7) {
      const int qn = min(QV, ((r0 + 4 * rq + 3) >> 5) + 1);
      if (QV == 4 && qn == 4) quad(std::integral_constant<int, QV>{}, rq);
      else if (QV == 4 && qn == 3) quad(std::integral_constant<int, (QV > 2 ? 3 : QV)>{}, rq);
      else if (qn == 2) quad(std::integral_constant<int, 2>{}, rq);
      else quad(std::integral_constant<int, 1>{}, rq);
    }
    BASE_STAMP(p, 3);
  }
  return fail_col;
#undef TT
}

template <int BN>
__global__ void __launch_bounds__(256, BN == 64 ? 2 : 1) potrf_trtri_base_kernel(
    const double* __restrict__ A, double* __restrict__ L, double* __restrict__ M, int ld, long long sbatch,
    double* __restrict__ logdet_part, int logdet_stride, int blk, int* __restrict__ info, int row_offset) {
  constexpr int BASE_N = BN, BASE_LD = BN + 1;
  extern __shared__ __align__(16) double sm[];
  double* T = sm;
  double* red = sm + BASE_N * BASE_LD + 2 * BASE_F + 64;
  const int tid = threadIdx.x, b = blockIdx.x;
  A += b * sbatch; L += b * sbatch; M += b * sbatch;
#define TT(i, k) T[(i) * BASE_LD + (k)]
  // block load: every element is an independent 8-byte cp.async (all in flight at once; a plain load loop
  // serialises 64 global-memory latencies per thread on the single resident CTA)
  for (int e = tid; e < BASE_N * BASE_N; e += 256) {
    const int i = e / BASE_N, k = e % BASE_N;
    if (k <= i) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&TT(i, k));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(A + (long long)i * ld + k));
      TT(k, i + 1) = (i == k) ? 1.0 : 0.0;
    }
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  const int fail_col = base_factor_smem<BN>(sm, tid);
  __syncthreads();
  for (int e = tid; e < BASE_N * BASE_N; e += 256) {
    const int i = e / BASE_N, k = e % BASE_N;
    L[(long long)i * ld + k] = (k <= i) ? TT(i, k) : 0.0;
    M[(long long)i * ld + k] = (k <= i) ? TT(k, i + 1) : 0.0;
  }
  // sum of log L_ii of this block, fixed order
  if (tid < 64) red[tid] = log(TT(tid, tid)) + (BN == 128 ? log(TT(tid + BN / 2, tid + BN / 2)) : 0.0);
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int k = 0; k < 64; k++) s += red[k];
    logdet_part[(long long)b * logdet_stride + blk] = s;
  }
  // every thread tracked the published fail columns identically; thread 0 reports
  if (tid == 0 && fail_col >= 0) atomicCAS(info + b, 0, row_offset + fail_col + 1);
#undef TT
}

// ------------------------------------------------------------------------------------------
// Gamma = M Y for a few right-hand sides (D <= 8): one warp per row, HBM-bound on M's lower part.
// Also emits rowsq[i] = sum_c Gamma[i][c]^2 for the quadratic form.
// Y is the user's unpadded [n][D]; Gamma is [np][D] (rows >= n are zero).
// ------------------------------------------------------------------------------------------
template <int DMAX>
__global__ void __launch_bounds__(256) trmv_lower_kernel(const double* __restrict__ M, int ld, long long sM,
                                                         const double* __restrict__ Y, int n, int D, long long sY,
                                                         double* __restrict__ Gm, long long sG,
                                                         double* __restrict__ rowsq, int np) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= np) return;
  const double* m = M + b * sM + (long long)row * ld;
  const double* y = Y + b * sY;
  double acc[DMAX];
#pragma unroll
  for (int c = 0; c < DMAX; c++) acc[c] = 0.0;
  const int kend = min(row + 1, n);
  int k = lane;
  for (; k + 96 < kend; k += 128) {            // 4 independent loads in flight per lane
    const double m0 = m[k], m1 = m[k + 32], m2 = m[k + 64], m3 = m[k + 96];
#pragma unroll
    for (int c = 0; c < DMAX; c++)
      if (c < D) {
        acc[c] = fma(m0, y[(long long)k * D + c], acc[c]);
        acc[c] = fma(m1, y[(long long)(k + 32) * D + c], acc[c]);
        acc[c] = fma(m2, y[(long long)(k + 64) * D + c], acc[c]);
        acc[c] = fma(m3, y[(long long)(k + 96) * D + c], acc[c]);
      }
  }
  for (; k < kend; k += 32) {
    const double mv = m[k];
#pragma unroll
    for (int c = 0; c < DMAX; c++)
      if (c < D) acc[c] = fma(mv, y[(long long)k * D + c], acc[c]);
  }
  double sq = 0.0;
#pragma unroll
  for (int c = 0; c < DMAX; c++) {
    double v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (c < D) {
      if (row >= n) v = 0.0;
      if (lane == 0) Gm[b * sG + (long long)row * D + c] = v;
      sq = fma(v, v, sq);
    }
  }
  if (lane == 0) rowsq[(long long)b * np + row] = sq;
}

// rowsq for the GEMM route (D > 8): Gamma padded [np][ldg]
__global__ void rowsq_kernel(const double* __restrict__ Gm, int ldg, long long sG, int Dp, double* __restrict__ rowsq, int np) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= np) return;
  const double* gr = Gm + b * sG + (long long)row * ldg;
  double s = 0.0;
  for (int c = lane; c < Dp; c += 32) { const double v = gr[c]; s = fma(v, v, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) rowsq[(long long)b * np + row] = s;
}

// ------------------------------------------------------------------------------------------
// out[c][q] = sum_{p in [p_lo(c), P)} Mat[p][c] * W[p][q]   for q < Q <= 8
//   lower=1: p_lo = first row of c's 32-column block (Mat lower-triangular, zeros above the diagonal)
//   square=1: W ignored, out[c][0] = sum_p Mat[p][c]^2
// 32 columns per CTA, 8 warps stride the rows, coalesced 256-byte row segments.
// ------------------------------------------------------------------------------------------
template <int QMAX>
__global__ void __launch_bounds__(256) colsum_weighted_kernel(const double* __restrict__ Mat, int ld, long long sMat, int P,
                                                              const double* __restrict__ W, int ldw, long long sW, int Q,
                                                              double* __restrict__ out, int ldo, long long sOut,
                                                              int ncols_out, int lower, int square) {
  __shared__ double part[8][32][QMAX + 1];
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * 32, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* mat = Mat + b * sMat;
  const double* w = W ? W + b * sW : nullptr;
  double acc[QMAX];
#pragma unroll
  for (int q = 0; q < QMAX; q++) acc[q] = 0.0;
  const int p_lo = lower ? c0 : 0;
  int pp = p_lo + warp;
  // 16 independent row loads in flight per warp (32 KB per CTA): the first column block of a lower-triangular Mat
  // streams all P rows through ONE CTA, which with 4 loads in flight was latency-bound at 5 GB/s (0.46 ms at
  // N = 8192, profiles/r01_metrics_c2_v1.txt)
  for (; pp + 120 < P; pp += 128) {
    double mv[16];
#pragma unroll
    for (int u = 0; u < 16; u++) mv[u] = mat[(long long)(pp + 8 * u) * ld + c0 + lane];
    if (square) {
#pragma unroll
      for (int u = 0; u < 16; u++) acc[0] = fma(mv[u], mv[u], acc[0]);
    } else {
#pragma unroll
      for (int u = 0; u < 16; u++)
#pragma unroll
        for (int q = 0; q < QMAX; q++)
          if (q < Q) acc[q] = fma(mv[u], w[(long long)(pp + 8 * u) * ldw + q], acc[q]);
    }
  }
  for (; pp + 24 < P; pp += 32) {               // 4 independent row loads in flight per warp
    double mv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) mv[u] = mat[(long long)(pp + 8 * u) * ld + c0 + lane];
    if (square) {
#pragma unroll
      for (int u = 0; u < 4; u++) acc[0] = fma(mv[u], mv[u], acc[0]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int q = 0; q < QMAX; q++)
          if (q < Q) acc[q] = fma(mv[u], w[(long long)(pp + 8 * u) * ldw + q], acc[q]);
    }
  }
  for (; pp < P; pp += 8) {
    const double mv = mat[(long long)pp * ld + c0 + lane];
    if (square) {
      acc[0] = fma(mv, mv, acc[0]);
    } else {
#pragma unroll
      for (int q = 0; q < QMAX; q++)
        if (q < Q) acc[q] = fma(mv, w[(long long)pp * ldw + q], acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < QMAX; q++) part[warp][lane][q] = acc[q];
  __syncthreads();
  const int nq = square ? 1 : Q;
  for (int e = threadIdx.x; e < 32 * nq; e += 256) {
    const int l = e / nq, q = e % nq;
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < 8; wv++) s += part[wv][l][q];
    if (c0 + l < ncols_out) out[b * sOut + (long long)(c0 + l) * ldo + q] = s;
  }
}

// nll[b] = 0.5 * sum_i rowsq[i] + D * sum_k logdet_part[k]      (fixed-order tree)
// ------------------------------------------------------------------------------------------
// alpha = S y from the LOWER triangle of S = Sigma^-1 (what the S = M^T M launch leaves in the workspace), one CTA per
// (problem, right-hand-side column).  Batched gradient evaluations compute S anyway; alpha = S y in ONE pass over
// S (8 N^2 / 2 bytes) replaces Gamma = M y and alpha = M^T Gamma (two passes over M by two kernels), and
// ||Gamma||^2 = y^T alpha gives the quadratic form: rowsq[i] = y_i alpha_i summed by nll_reduce_kernel as before.
// 32 x 32 tiles of the lower triangle, round-robin over the 8 warps; a lane holds column l of the tile (32 independent
// coalesced 256-byte row loads in flight), the tile's contribution to alpha[rows] is a warp reduction per row, the
// mirrored contribution to alpha[cols] is lane-local.  Per-warp partial vectors in shared memory, summed in fixed order.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) symv_lower_kernel(const double* __restrict__ S, int ld, long long sS, int n, int np,
                                                         const double* __restrict__ y, int D, long long sy,
                                                         double* __restrict__ alpha, long long salpha,
                                                         double* __restrict__ rowsq) {
  extern __shared__ __align__(16) double sv_sm[];            // [8][np] per-warp partial alpha, then [np] y column
  const int b = blockIdx.x, col = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* part = sv_sm + (size_t)warp * np;
  double* ysh = sv_sm + (size_t)8 * np;
  const double* Sb = S + (long long)b * sS;
  for (int i = tid; i < np; i += 256) ysh[i] = (i < n) ? y[(long long)b * sy + (long long)i * D + col] : 0.0;
  for (int i = lane; i < np; i += 32) part[i] = 0.0;
  __syncthreads();
  const int nt = (n + 31) >> 5;                              // tile rows / columns that hold data
  const int ntiles = nt * (nt + 1) / 2;
  for (int t = warp; t < ntiles; t += 8) {
    int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    const int tj = t - ti * (ti + 1) / 2;
    const int i0 = ti * 32, j0 = tj * 32;
    const bool diag = ti == tj;
    double v[32];
#pragma unroll
    for (int r = 0; r < 32; r++) v[r] = Sb[(long long)(i0 + r) * ld + j0 + lane];
    const double yc = ysh[j0 + lane];
    double colacc = 0.0;
#pragma unroll
    for (int r = 0; r < 32; r++) {
      // element (i0 + r, j0 + lane): on a diagonal tile only r >= lane exists; its mirror (r > lane) feeds alpha[j0 + lane]
      const bool lower = !diag || r >= lane;
      double rowc = lower ? v[r] * yc : 0.0;
      if (lower && (!diag || r > lane)) colacc = fma(v[r], ysh[i0 + r], colacc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rowc += __shfl_xor_sync(0xffffffffu, rowc, o);
      if (lane == 0) part[i0 + r] += rowc;
    }
    __syncwarp();
    part[j0 + lane] += colacc;
    __syncwarp();
  }
  __syncthreads();
  for (int i = tid; i < np; i += 256) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) a += sv_sm[(size_t)w * np + i];
    if (i >= n) a = 0.0;
    alpha[(long long)b * salpha + (long long)i * D + col] = a;
    if (D == 1) rowsq[(long long)b * np + i] = a * ysh[i];
  }
}

// rowsq[i] = sum_c y[i][c] alpha[i][c] for D > 1 (the per-column symv launches above cannot combine columns)
__global__ void __launch_bounds__(256) rowdot_kernel(const double* __restrict__ y, long long sy, const double* __restrict__ alpha,
                                                     long long salpha, int n, int np, int D, double* __restrict__ rowsq) {
  const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= np) return;
  double a = 0.0;
  if (i < n)
    for (int c = 0; c < D; c++) a = fma(y[(long long)b * sy + (long long)i * D + c], alpha[(long long)b * salpha + (long long)i * D + c], a);
  rowsq[(long long)b * np + i] = a;
}

__global__ void __launch_bounds__(256) nll_reduce_kernel(const double* __restrict__ rowsq, int np,
                                                         const double* __restrict__ logdet_part, int nblk, int D,
                                                         double* __restrict__ nll, double* __restrict__ logdet_out) {
  __shared__ double s1[256], s2[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  double a = 0.0, l = 0.0;
  for (int i = tid; i < np; i += 256) a += rowsq[(long long)b * np + i];
  for (int i = tid; i < nblk; i += 256) l += logdet_part[(long long)b * nblk + i];
  s1[tid] = a; s2[tid] = l;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { s1[tid] += s1[tid + o]; s2[tid] += s2[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    if (nll) nll[b] = 0.5 * s1[0] + (double)D * s2[0];
    if (logdet_out) logdet_out[b] = 2.0 * s2[0];
  }
}

// ------------------------------------------------------------------------------------------
// Gradient contraction.  With G = dNLL/dSigma (= 0.5 (D S - alpha alpha^T), S = Sigma^-1) and
// K_ij = amp exp(-0.5 sum_k dz_ijk^2), dz = (x_i - x_j) o w:
//   dNLL/dw_k  = -(1/w_k) sum_ij G_ij K_ij dz_ijk^2,  dNLL/damp = (1/amp) sum_ij G_ij K_ij,
//   dNLL/d diag_add_i = G_ii.
// One CTA per 64x64 lower tile of G (mirror counted twice), K re-evaluated from x (cheaper than
// an N^2 read); per-CTA partials are reduced in fixed order by grad_finish_kernel.
// src holds S (src_is_G = 0: G formed here from alpha, D <= 8) or G itself (src_is_G = 1).
// ------------------------------------------------------------------------------------------
constexpr int GRAD_T = 64;
constexpr int GRAD_DMAX = 128;  // input dims held in shared memory per tile (gen-2023 NAR feeds [x, y_low]: d = 2 + 64 at the C1 size)

struct GradParams {
  const double* src; int ld; long long ssrc;
  const double* x; int n, d; long long sx;
  const double* w; long long sw;
  const double* amp; long long samp;
  const double* alpha; int D; long long salpha;   // [n][D] (unpadded), used when !src_is_G
  int src_is_G;
  double* partial; int npart;                     // [batch][npart][d+1]
  double* g_diag; long long sgd;                  // [batch][n] or NULL
  double* G_out; long long sGo;                   // optional full [n][n] dNLL/dSigma (symmetric), or NULL
  int have_k;                                     // 0: no kernel part (covariance-input mode)
};

// Shared-memory plan of grad_contract_kernel for input dimension d (doubles).  The scaled inputs of the tile's rows
// and columns are held with the dimension count padded to whole 8-wide MMA fragments.
__host__ __device__ inline int grad_nz(int d) { return (d + 7) / 8; }               // 8-wide fragments over the dimensions
__host__ __device__ inline int grad_ldz(int d) { return grad_nz(d) * 8 + 4; }       // row stride of zi / zj (conflict-free fragments)
constexpr int GRAD_LDW = GRAD_T + 4;
__host__ __device__ inline size_t grad_smem_doubles(int d) {
  return (size_t)2 * GRAD_T * grad_ldz(d) + (size_t)GRAD_T * GRAD_LDW + 2 * GRAD_T * 8 + 8 * GRAD_T + 8 * (GRAD_DMAX + 1);
}

// Formulation (everything level-3 on the FP64 tensor pipe, the elementwise part is one exp per pair):
//   z = (x - x_0) o w   (the kernel is translation invariant; centring on the problem's first point keeps the
//                        norm expansions below well conditioned when the inputs carry a large offset)
//   sq_ij = |z_i|^2 + |z_j|^2 - 2 (Zi Zj^T)_ij                      cross term: DMMA, k = d
//   W_ij  = G_ij exp(-sq_ij / 2)   (x2 below the diagonal: the mirrored pair; diagonal: G_ii)
//   sum_ij W_ij (z_ik - z_jk)^2 = sum_i R_i z_ik^2 + sum_j C_j z_jk^2 - 2 sum_i z_ik (W Zj)_ik     W Zj: DMMA, k = 64
// with R / C the row / column sums of the tile's W.  The previous version evaluated (z_ik - z_jk)^2 pair by pair and
// dimension by dimension twice (5 d + 45 FP64 instructions per pair, FP64-FMA bound: 0.88 ms per 512 problems of
// N = 512, d = 8, profiles/r01_launches_c5_v6.csv); this one needs about 35 per pair independent of d.
// GOUT: also write the full symmetric dNLL/dSigma (covariance-input mode); kept out of the common instantiation.
#ifndef GRAD_MIN_CTAS
#define GRAD_MIN_CTAS 3
#endif
template <bool GOUT>
__global__ void __launch_bounds__(256, GRAD_MIN_CTAS) grad_contract_kernel(const GradParams p) {
  extern __shared__ __align__(16) double gsm[];
  const int b = blockIdx.y;
  int t = blockIdx.x;
  int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (ti * (ti + 1) / 2 > t) --ti;
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  const int tj = t - ti * (ti + 1) / 2;
  const int i0 = ti * GRAD_T, j0 = tj * GRAD_T;
  const int d = p.have_k ? p.d : 0;
  const int nz = grad_nz(d), ldz = grad_ldz(p.d), dz = nz * 8;       // dz: dimensions padded with zeros
  double* zi = gsm;                         // [64][ldz]
  double* zj = zi + GRAD_T * ldz;           // [64][ldz]
  double* Wsm = zj + GRAD_T * ldz;          // [64][GRAD_LDW]
  double* ai = Wsm + GRAD_T * GRAD_LDW;     // [64][8]
  double* aj = ai + GRAD_T * 8;             // [64][8]
  double* nrm = aj + GRAD_T * 8;            // [2][64] squared norms of the rows of zi, zj
  double* rowp = nrm + 2 * GRAD_T;          // [4 warp columns][64] partial row sums of W
  double* colp = rowp + 4 * GRAD_T;         // [2 warp rows][64] partial column sums of W
  double* red = colp + 2 * GRAD_T;          // [8 warps][GRAD_DMAX + 1]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const double* x = p.x ? p.x + b * p.sx : nullptr;
  const double* w = p.w ? p.w + b * p.sw : nullptr;
  const int wm = warp >> 2, wn = warp & 3;           // 2 x 4 warps, warp tile 32 x 16
  // the thread's 16 entries of S (or G), requested before anything else: their latency hides behind the z staging
  double2 sreg[4][2];
  {
    const double* src = p.src + b * p.ssrc;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int gi = i0 + wm * 32 + i * 8 + g;
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int gc = j0 + wn * 16 + j * 8 + tq * 2;
        sreg[i][j] = (gi < p.n && gc <= gi) ? *reinterpret_cast<const double2*>(src + (long long)gi * p.ld + gc)
                                           : make_double2(0.0, 0.0);
      }
    }
  }
  if (d > 0) {          // 4 threads per row, strided over the (padded) dimensions: no index divisions
    const int r = tid >> 2, gi = i0 + r, gj = j0 + r;
    for (int k = tid & 3; k < dz; k += 4) {
      const bool kin = k < d;
      const double x0 = kin ? x[k] : 0.0, wk = kin ? w[k] : 0.0;
      zi[r * ldz + k] = (kin && gi < p.n) ? (x[(long long)gi * d + k] - x0) * wk : 0.0;
      zj[r * ldz + k] = (kin && gj < p.n) ? (x[(long long)gj * d + k] - x0) * wk : 0.0;
    }
  }
  if (!p.src_is_G) {
    const double* al = p.alpha + b * p.salpha;
    const int r = tid >> 2;
    for (int c = tid & 3; c < 8; c += 4) {
      ai[r * 8 + c] = (i0 + r < p.n && c < p.D) ? al[(long long)(i0 + r) * p.D + c] : 0.0;
      aj[r * 8 + c] = (j0 + r < p.n && c < p.D) ? al[(long long)(j0 + r) * p.D + c] : 0.0;
    }
  }
  __syncthreads();
  if (tid < 2 * GRAD_T) {
    const double* zr = (tid < GRAD_T ? zi : zj) + (tid % GRAD_T) * ldz;
    double s = 0.0;
    for (int k = 0; k < dz; k++) s = fma(zr[k], zr[k], s);
    nrm[tid] = s;
  }
  // ---- cross term Zi Zj^T ----
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  for (int kk = 0; kk < dz; kk += 4) {
    double af[4], bf[2];
#pragma unroll
    for (int i = 0; i < 4; i++) af[i] = zi[(wm * 32 + i * 8 + g) * ldz + kk + tq];
#pragma unroll
    for (int j = 0; j < 2; j++) bf[j] = zj[(wn * 16 + j * 8 + g) * ldz + kk + tq];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
  __syncthreads();                                   // norms
  // ---- W = G o K/amp on the thread's 16 pairs (C-fragment layout: row g, columns 2 tq, 2 tq + 1) ----
  // Straight-line over the 16 pairs (select, no branches): the 16 exp() polynomial chains interleave.  A branchy
  // version left the FP64 pipe 14 % busy, stalled on the dependent chain of one exp at a time
  // (profiles/r01_ncu_grad_contract_v2.txt).
  const double amp = p.have_k ? p.amp[b * p.samp] : 0.0;
  double wv[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = wm * 32 + i * 8 + g;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int c0 = wn * 16 + j * 8 + tq * 2;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const double sq = fmax(nrm[r] + nrm[GRAD_T + c0 + e] - 2.0 * acc[i][j][e], 0.0);
        wv[i][j][e] = exp_nonpos(-0.5 * sq);
      }
    }
  }
  double sumW = 0.0;
  double rsum[4] = {0.0, 0.0, 0.0, 0.0}, csum[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = wm * 32 + i * 8 + g, gi = i0 + r;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int c0 = wn * 16 + j * 8 + tq * 2;
      const double sv[2] = {sreg[i][j].x, sreg[i][j].y};
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = c0 + e, gj = j0 + c;
        const bool valid = gi < p.n && gj <= gi;
        double G = sv[e];
        if (!p.src_is_G) {
          double aa = ai[r * 8] * aj[c * 8];
          if (p.D > 1)
            for (int cc = 1; cc < p.D; cc++) aa = fma(ai[r * 8 + cc], aj[c * 8 + cc], aa);
          G = 0.5 * ((double)p.D * G - aa);
        }
        if (GOUT && valid) {
          double* go = p.G_out + b * p.sGo;
          go[(long long)gi * p.n + gj] = G;
          go[(long long)gj * p.n + gi] = G;
        }
        if (valid && gi == gj && p.g_diag) p.g_diag[b * p.sgd + gi] = G;
        // diagonal: K_ii / amp = 1, dz = 0; below it the pair stands for (i,j) and (j,i); amplitude applied at the end
        const double wgt = !valid ? 0.0 : (gi == gj ? G : (p.have_k ? 2.0 * G * wv[i][j][e] : 0.0));
        wv[i][j][e] = wgt;
      }
      sumW += wv[i][j][0] + wv[i][j][1];             // sum G o (K / amp): d/d amp needs no division (amp may be 0)
      rsum[i] += wv[i][j][0] + wv[i][j][1];
      csum[j][0] += wv[i][j][0];
      csum[j][1] += wv[i][j][1];
      *reinterpret_cast<double2*>(Wsm + r * GRAD_LDW + c0) = make_double2(wv[i][j][0], wv[i][j][1]);
    }
  }
  // row sums: over the 4 lanes of a row (tq); column sums: over the 8 lanes of a column (g) - fixed order
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double v = rsum[i];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (tq == 0) rowp[wn * GRAD_T + wm * 32 + i * 8 + g] = v;
  }
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      double v = csum[j][e];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (g == 0) colp[wm * GRAD_T + wn * 16 + j * 8 + tq * 2 + e] = v;
    }
  __syncthreads();
  // ---- P = W Zj: warp owns rows warp * 8 .. + 8; per 8-dimension fragment one 64-deep DMMA chain ----
  // contribution of element (row, k): zi[row][k] (R_row zi[row][k] - 2 P[row][k]) + C_row zj[row][k]^2
  if (d > 0) {
    const int row = warp * 8 + g;
    const double Rr = (rowp[row] + rowp[GRAD_T + row]) + (rowp[2 * GRAD_T + row] + rowp[3 * GRAD_T + row]);
    const double Cr = colp[row] + colp[GRAD_T + row];
    for (int nf = 0; nf < nz; nf++) {
      double p0 = 0.0, p1 = 0.0;
#pragma unroll 4
      for (int kk = 0; kk < GRAD_T; kk += 4)
        dmma884(p0, p1, Wsm[row * GRAD_LDW + kk + tq], zj[(kk + tq) * ldz + nf * 8 + g]);
      const int k0 = nf * 8 + tq * 2;
      const double2 zi2 = *reinterpret_cast<const double2*>(zi + row * ldz + k0);
      const double2 zj2 = *reinterpret_cast<const double2*>(zj + row * ldz + k0);
      double v0 = fma(zi2.x, fma(Rr, zi2.x, -2.0 * p0), Cr * zj2.x * zj2.x);
      double v1 = fma(zi2.y, fma(Rr, zi2.y, -2.0 * p1), Cr * zj2.y * zj2.y);
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {             // sum over the 8 rows of the fragment (lanes with equal tq)
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      }
      if (g == 0) {
        if (k0 < d) red[warp * (GRAD_DMAX + 1) + k0] = v0 * amp;
        if (k0 + 1 < d) red[warp * (GRAD_DMAX + 1) + k0 + 1] = v1 * amp;
      }
    }
  }
  {
    double v = sumW;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * (GRAD_DMAX + 1) + GRAD_DMAX] = v;
  }
  __syncthreads();
  double* part = p.partial + ((long long)b * p.npart + t) * (p.d + 1);
  for (int k = tid; k <= p.d; k += 256) {
    const int kk = (k == p.d) ? GRAD_DMAX : k;
    double s = 0.0;
    if (k == p.d || p.have_k)
      for (int wv = 0; wv < 8; wv++) s += red[wv * (GRAD_DMAX + 1) + kk];
    part[k] = s;
  }
}

// g_w[k] = -(1/w_k) sum_t partial[t][k];  g_amp = sum_t partial[t][d]  (partials of G o K/amp).
// grid (d+1, batch): each CTA reduces one component over all tiles with a fixed-order tree.
__global__ void __launch_bounds__(256) grad_finish_kernel(const double* __restrict__ partial, int npart, int d,
                                                          const double* __restrict__ w, long long sw,
                                                          const double* __restrict__ amp, long long samp,
                                                          double* __restrict__ g_w, double* __restrict__ g_amp) {
  __shared__ double red[256];
  const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const double* pp = partial + (long long)b * npart * (d + 1) + k;
  double s = 0.0;
  for (int t = tid; t < npart; t += 256) s += pp[(long long)t * (d + 1)];
  red[tid] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    if (k < d) g_w[(long long)b * d + k] = -red[0] / w[b * sw + k];
    else g_amp[b] = red[0];
  }
}

// dst[r][c] = (r < rows_src && c < cols_src) ? src[r][c] : 0   over [rows_dst][cols_dst]
__global__ void pad_copy_kernel(const double* __restrict__ src, int rows_src, int cols_src, int lds, long long ssrc,
                                double* __restrict__ dst, int rows_dst, int cols_dst, int ldd, long long sdst) {
  const int b = blockIdx.z;
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= rows_dst || c >= cols_dst) return;
  double v = 0.0;
  if (r < rows_src && c < cols_src) v = src[b * ssrc + (long long)r * lds + c];
  dst[b * sdst + (long long)r * ldd + c] = v;
}

// dst[i][j] = a * src[i][j] + add[b]   (cov = Kxx - V^T V + noise offset is produced by the GEMM epilogue;
// this kernel only copies the padded result into the user's [ns][ns] array)
__global__ void unpad_copy_kernel(const double* __restrict__ src, int lds, long long ssrc,
                                  double* __restrict__ dst, int rows, int cols, long long sdst, int lower_only) {
  const int b = blockIdx.z;
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (r >= rows || c >= cols) return;
  dst[b * sdst + (long long)r * cols + c] = (lower_only && c > r) ? 0.0 : src[b * ssrc + (long long)r * lds + c];
}

// var[s] = kss_amp + offset - colsq[s]   (diagonal predictive variance; K(x*,x*)_ss = amp)
__global__ void var_diag_kernel(const double* __restrict__ colsq, const double* __restrict__ amp, long long samp,
                                const double* __restrict__ offset, long long soff, double* __restrict__ var, int ns,
                                int ld_colsq) {
  const int b = blockIdx.y, s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  var[(long long)b * ns + s] = amp[b * samp] + (offset ? offset[b * soff] : 0.0) - colsq[(long long)b * ld_colsq + s];
}

}  // namespace ffgp
