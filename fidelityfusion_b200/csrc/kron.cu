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
#include <atomic>
#include <cstring>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/ffgp.h"
#include "gemm_dmma.cuh"
#include "syevj_hestenes.cuh"

namespace ffgp {
int fail(int code, const char* fmt, const char* a);
extern std::atomic<unsigned long long> g_launches;

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
// Streaming n-mode product for small factor matrices (I, J <= 128: every output mode of the reference's configs
// and the N-mode of C4).  History (profiles/r01_kron_bench_v{1,2}.txt): the generic kernel above reached 10 % of HBM
// (factor reloaded per chunk, scalar loads, div/mod per element, no overlap); a CTA-synchronous cp.async ring
// reached 41 % at 134 MB (one __syncthreads and three 64-bit divisions per 16 KB chunk left both the DMMA pipe and
// HBM waiting).  This version has NO CTA-level synchronisation in the loop:
//   * the factor matrix is staged ONCE per CTA in shared memory as ms[j][k] (stride SA = 4 mod 16 doubles: the
//     4 rows x 4 k of a half-warp fragment load hit 16 distinct bank pairs); grid.y splits J in blocks of 64;
//   * every WARP streams its own tiles through a private 4-stage cp.async ring (16-byte copies) and synchronises
//     with __syncwarp only; tiles are dealt round-robin over all warps of the grid, so neighbouring warps touch
//     neighbouring 64-byte segments at the same time:
//       LAST == false (inner even): tile = 32 k-rows x 8 generalised columns (row stride 12 doubles),
//                                   C[j][col]: the warp loops the J/8 m-tiles;
//       LAST == true  (inner == 1): tile = 8 rows x 32 k (row stride 36),  C[row][j]: loops the J/8 n-tiles;
//   * generalised column -> (outer, inner) indices advance incrementally (no division in the loop);
//   * results leave as 16-byte stores.
// Algorithmic bytes: 8 (I + J) per generalised column + the factor.
// ---------------------------------------------------------------------------------------------
constexpr int MS_KB = 32, MS_STAGES = 4;
constexpr int MS_STAGE_DOUBLES = 256;     // 32 k x 8 columns, or 8 rows x 32 k: 2 KB, XOR-swizzled instead of padded
constexpr int MS_WARPS = 8;

struct ModeSmallParams {
  const double* t; const double* mat; double* out;
  long long ncols, inner;
  int I, J, transpose_mat, Ip, Jp, SA;
};

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// Shared-memory tile layouts (doubles), both conflict-free for the m8n8k4 fragment loads without padding:
//   LAST == false: element (k, col) at k * 8 + (col ^ (4 * ((k >> 1) & 1)))
//   LAST == true : element (row, k) at row * 32 + (k ^ (4 * (row & 3)))
// MT = number of 8-row tiles of the factor block handled per warp (J block padded to MT * 8 rows of zeros).
// (ncu of the previous version, profiles/r01_ncu_mode_dot_v3.txt: 800 warp instructions per 2 KB tile - runtime-
//  predicated 8 x 8 unrolled loops - kept the issue slots 47 % busy with 4 warps per scheduler and DRAM at 27 %.)
template <bool LAST, int MT>
__global__ void __launch_bounds__(256, 3) mode_dot_small_kernel(const ModeSmallParams p) {
  extern __shared__ __align__(16) double msm[];
  double* ms = msm;                                         // [MT * 8 rows of this CTA's J block][SA]
  const int j0 = blockIdx.y * (MT * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  double* stages = msm + (size_t)(MT * 8) * p.SA + (size_t)warp * MS_STAGES * MS_STAGE_DOUBLES;   // warp-private ring
  const int I = p.I, J = p.J, SA = p.SA;
  const long long inner = p.inner, ncols = p.ncols;
  for (int e = tid; e < MT * 8 * SA; e += 256) {
    const int jl = e / SA, k = e - jl * SA, j = j0 + jl;
    double v = 0.0;
    if (j < J && k < I) v = p.transpose_mat ? p.mat[(long long)k * J + j] : p.mat[(long long)j * I + k];
    ms[e] = v;
  }
  __syncthreads();                                          // the only CTA barrier
  const int nkb = (p.Ip + MS_KB - 1) / MS_KB;
  const long long nblocks = (ncols + 7) >> 3;               // 8-column (LAST: 8-row) blocks
  const long long wstride = (long long)gridDim.x * MS_WARPS;
  const long long wb0 = (long long)blockIdx.x * MS_WARPS + warp;
  const int my_blocks = (nblocks > wb0) ? (int)((nblocks - wb0 + wstride - 1) / wstride) : 0;
  const int nq = my_blocks * nkb;
  // per-lane generalised column of the LOAD side (advances by wstride * 8 per block) and of the STORE side
  const long long dq = (wstride * 8) / inner, dr = (wstride * 8) - dq * inner;
  long long lc = wb0 * 8 + (LAST ? 0 : 2 * tq), lo = LAST ? 0 : lc / inner, lci = LAST ? 0 : lc - lo * inner;   // load cursor
  long long sc = lc, so = lo, sci = lci;                                                                        // store cursor
  int lkb = 0, issued = 0;

  auto issue = [&]() {
    if (issued < nq) {
      double* st = stages + (issued & (MS_STAGES - 1)) * MS_STAGE_DOUBLES;
      if (!LAST) {
        const bool cok = lc < ncols;
        const double* src = p.t + (lo * I + lkb * MS_KB) * inner + lci;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int kl = g + 8 * i;
          double* dst = st + kl * 8 + ((2 * tq) ^ (4 * ((kl >> 1) & 1)));
          if (cok && lkb * MS_KB + kl < I) cp_async16(dst, src + (long long)kl * inner);
          else { dst[0] = 0.0; dst[1] = 0.0; }
        }
      } else {
        const int kv = lane & 15, k = lkb * MS_KB + 2 * kv;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int r = (lane >> 4) + 2 * i;
          double* dst = st + r * 32 + ((2 * kv) ^ (4 * (r & 3)));
          if (lc + r < ncols && k < I) cp_async16(dst, p.t + (lc + r) * I + k);
          else { dst[0] = 0.0; dst[1] = 0.0; }
        }
      }
      ++issued;
      if (++lkb == nkb) {
        lkb = 0;
        lc += wstride * 8;
        if (!LAST) { lo += dq; lci += dr; if (lci >= inner) { lci -= inner; ++lo; } }
      }
    }
    cp_async_commit();
  };

  for (int q = 0; q < MS_STAGES - 1; q++) issue();
  double acc[MT][2];
#pragma unroll
  for (int i = 0; i < MT; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  int kb = 0;
  const uint32_t ms_u32 = (uint32_t)__cvta_generic_to_shared(ms) + (uint32_t)((g * SA + tq) * 8);
  const uint32_t st_u32 = (uint32_t)__cvta_generic_to_shared(stages) +
                          (uint32_t)((LAST ? g * 32 + tq : tq * 8 + (g ^ (4 * (tq >> 1)))) * 8);
  const uint32_t m_stride = (uint32_t)(SA * 64);
  const uint32_t a_swz = (uint32_t)(g & 3);
  for (int q = 0; q < nq; q++) {
    cp_async_wait<MS_STAGES - 2>();
    __syncwarp();                             // tile q landed for every lane; the stage of tile q-1 is free
    issue();
    const uint32_t s_base = st_u32 + (uint32_t)((q & (MS_STAGES - 1)) * MS_STAGE_DOUBLES * 8);
    const uint32_t m_base = ms_u32 + (uint32_t)(kb * MS_KB * 8);
    const int ksteps = min(MS_KB, p.Ip - kb * MS_KB) >> 2;
#pragma unroll 4
    for (int ks = 0; ks < ksteps; ks++) {
      const double sf = lds_f64(LAST ? s_base + (((uint32_t)ks ^ a_swz) << 5) : s_base + (uint32_t)(ks << 8));
#pragma unroll
      for (int i = 0; i < MT; i++) {
        const double mf = lds_f64(m_base + ks * 32 + i * m_stride);
        if (LAST) dmma884(acc[i][0], acc[i][1], sf, mf);
        else dmma884(acc[i][0], acc[i][1], mf, sf);
      }
    }
    if (++kb == nkb) {
      kb = 0;
      if (!LAST) {
        if (sc < ncols) {
          double* dst = p.out + (so * J + j0 + g) * inner + sci;
#pragma unroll
          for (int i = 0; i < MT; i++)
            if (j0 + i * 8 + g < J) *reinterpret_cast<double2*>(dst + (long long)(i * 8) * inner) = make_double2(acc[i][0], acc[i][1]);
        }
        sc += wstride * 8; so += dq; sci += dr;
        if (sci >= inner) { sci -= inner; ++so; }
      } else {
        const long long r = sc + g;
        if (r < ncols) {
          double* dst = p.out + r * J + j0 + tq * 2;
#pragma unroll
          for (int i = 0; i < MT; i++) {
            const int j = j0 + i * 8 + tq * 2;
            if (!(J & 1) && j + 1 < J) *reinterpret_cast<double2*>(dst + i * 8) = make_double2(acc[i][0], acc[i][1]);
            else {
              if (j < J) dst[i * 8] = acc[i][0];
              if (j + 1 < J) dst[i * 8 + 1] = acc[i][1];
            }
          }
        }
        sc += wstride * 8;
      }
#pragma unroll
      for (int i = 0; i < MT; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    }
  }
  cp_async_wait<0>();
}

template <bool LAST, int MT>
static cudaError_t launch_mode_small(const ModeSmallParams& p, int sms, cudaStream_t st) {
  auto kern = mode_dot_small_kernel<LAST, MT>;
  const size_t smem = ((size_t)(MT * 8) * p.SA + (size_t)MS_WARPS * MS_STAGES * MS_STAGE_DOUBLES) * sizeof(double);
  static PerDeviceOnce once;
  bool& attr = *once.slot();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int jblocks = (p.Jp + MT * 8 - 1) / (MT * 8);
  const long long nchunks = (p.ncols + 63) / 64;              // 8 warps x 8-column blocks per CTA pass
  const int per_sm = std::max(1, std::min(3, (int)((224 * 1024) / (smem + 1024))));
  const int gx = (int)std::min<long long>(nchunks, std::max(1, sms * per_sm / jblocks));
  kern<<<dim3(gx, jblocks), 256, smem, st>>>(p);
  return cudaGetLastError();
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

// ---------------------------------------------------------------------------------------------
// Streaming mode Gram (replaces mode_gram_kernel for 16-byte-alignable layouts; that kernel reached 3-7 % of HBM and
// a CTA-synchronous cp.async version 50 %, profiles/r01_kron_bench_v{1,2}.txt).  Same structure as
// mode_dot_small_kernel: no CTA barrier in the loop, every warp streams its own 8-column blocks of X and Y through a
// private 4-stage cp.async ring and accumulates a whole (8 MT) x (8 MT) tile of G in registers; grid.z / grid.y pick
// the 32 x 32 tile of G, grid.x the interleaved slice of the column range.  Warps are summed in fixed order at the
// end and slices in index order by gram_reduce_kernel (deterministic, no atomics).
// Stage = X tile | Y tile, 256 doubles each, XOR-swizzled:
//   INNER1 == false (inner even): element (row r, col c)  at r * 8  + (c ^ (4 * ((r >> 1) & 1)))
//   INNER1 == true  (inner == 1): element (col c, row a)  at c * 32 + (a ^ (4 * (c & 3)))
// ---------------------------------------------------------------------------------------------
constexpr int GS_STAGES = 3, GS_STAGE_DOUBLES = 512, GS_WARPS = 8;   // 12 KB per warp, 96 KB per CTA: 2 CTAs / SM

template <bool INNER1, int MT>
__global__ void __launch_bounds__(256, 2) mode_gram_stream_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                                                                  double* __restrict__ part, long long ncols, long long inner,
                                                                  int Ja, int Jb) {
  extern __shared__ __align__(16) double gsm2[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int a0 = blockIdx.z * 32, b0 = blockIdx.y * 32;
  double* stages = gsm2 + (size_t)warp * GS_STAGES * GS_STAGE_DOUBLES;
  const long long nblocks = (ncols + 7) >> 3;
  const long long wstride = (long long)gridDim.x * GS_WARPS;
  const long long wb0 = (long long)blockIdx.x * GS_WARPS + warp;
  const int nq = (nblocks > wb0) ? (int)((nblocks - wb0 + wstride - 1) / wstride) : 0;
  const long long dq = (wstride * 8) / inner, dr = (wstride * 8) - dq * inner;
  long long lc = wb0 * 8 + (INNER1 ? 0 : 2 * tq), lo = INNER1 ? 0 : lc / inner, lci = INNER1 ? 0 : lc - lo * inner;
  int issued = 0;

  auto issue = [&]() {
    if (issued < nq) {
      double* st = stages + (issued % GS_STAGES) * GS_STAGE_DOUBLES;
      if (!INNER1) {
        const bool cok = lc < ncols;
#pragma unroll
        for (int i = 0; i < MT; i++) {
          const int r = g + 8 * i;                                           // row of the 8 MT-row tile
          const int swz = (2 * tq) ^ (4 * ((r >> 1) & 1));
          double* dx = st + r * 8 + swz;
          double* dy = dx + 256;
          if (cok && a0 + r < Ja) cp_async16(dx, X + (lo * Ja + a0 + r) * inner + lci);
          else { dx[0] = 0.0; dx[1] = 0.0; }
          if (cok && b0 + r < Jb) cp_async16(dy, Y + (lo * Jb + b0 + r) * inner + lci);
          else { dy[0] = 0.0; dy[1] = 0.0; }
        }
      } else {
        // 8 columns x (8 MT) rows per operand: 4 MT 16-byte vectors per column
#pragma unroll
        for (int i = 0; i < MT; i++) {
          const int v = lane + 32 * i;
          const int c = v / (4 * MT), e = 2 * (v % (4 * MT));
          const int off = c * 32 + (e ^ (4 * (c & 3)));
          double* dx = st + off;
          double* dy = dx + 256;
          if (lc + c < ncols && a0 + e < Ja) cp_async16(dx, X + (lc + c) * Ja + a0 + e);
          else { dx[0] = 0.0; dx[1] = 0.0; }
          if (lc + c < ncols && b0 + e < Jb) cp_async16(dy, Y + (lc + c) * Jb + b0 + e);
          else { dy[0] = 0.0; dy[1] = 0.0; }
        }
      }
      ++issued;
      lc += wstride * 8;
      if (!INNER1) { lo += dq; lci += dr; if (lci >= inner) { lci -= inner; ++lo; } }
    }
    cp_async_commit();
  };

  for (int q = 0; q < GS_STAGES - 1; q++) issue();
  double acc[MT][MT][2];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < MT; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  for (int q = 0; q < nq; q++) {
    cp_async_wait<GS_STAGES - 2>();
    __syncwarp();
    issue();
    const double* st = stages + (q % GS_STAGES) * GS_STAGE_DOUBLES;
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
      double af[MT], bf[MT];
#pragma unroll
      for (int i = 0; i < MT; i++) {
        if (!INNER1) {
          const int r = i * 8 + g, c = ks * 4 + tq;
          const int off = r * 8 + (c ^ (4 * ((r >> 1) & 1)));
          af[i] = st[off];
          bf[i] = st[256 + off];
        } else {
          const int c = ks * 4 + tq, r = i * 8 + g;
          const int off = c * 32 + (r ^ (4 * (c & 3)));
          af[i] = st[off];
          bf[i] = st[256 + off];
        }
      }
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < MT; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // cross-warp sum in fixed order through shared memory: red[warp][32][34]
  double* red = gsm2;
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < MT; j++) {
      double* d = red + (size_t)warp * 32 * 34 + (i * 8 + g) * 34 + j * 8 + tq * 2;
      d[0] = acc[i][j][0];
      d[1] = acc[i][j][1];
    }
  __syncthreads();
  double* dst = part + (long long)blockIdx.x * Ja * Jb;
  for (int e = tid; e < (8 * MT) * (8 * MT); e += 256) {
    const int r = e / (8 * MT), c = e - r * (8 * MT);
    const int a = a0 + r, b = b0 + c;
    if (a < Ja && b < Jb) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < GS_WARPS; k++) v += red[(size_t)k * 32 * 34 + r * 34 + c];
      dst[(long long)a * Jb + b] = v;
    }
  }
}

template <bool INNER1, int MT>
static cudaError_t launch_gram_stream(const double* X, const double* Y, double* part, long long ncols, long long inner, int Ja,
                                      int Jb, int ns, cudaStream_t st) {
  const size_t smem = std::max((size_t)GS_WARPS * GS_STAGES * GS_STAGE_DOUBLES, (size_t)GS_WARPS * 32 * 34) * sizeof(double);
  auto kern = mode_gram_stream_kernel<INNER1, MT>;
  static PerDeviceOnce once;
  bool& attr = *once.slot();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  kern<<<dim3(ns, (Jb + 31) / 32, (Ja + 31) / 32), 256, smem, st>>>(X, Y, part, ncols, inner, Ja, Jb);
  return cudaGetLastError();
}

// out[i] = sum_z part[z][i], fixed order: 32 elements x 8 slice lanes per block - lane q adds the slices z = q, q + 8, ...
// in order, the 8 lane sums are added in lane order.  (One thread per element walking all slices serially took 27 us on
// the 32 x 32 Gram matrices of the C4 step: ~270 dependent L2 round trips for 4 blocks of work.)
__global__ void __launch_bounds__(256) gram_reduce_kernel(const double* __restrict__ part, int nslices, long long nelem,
                                                          double* __restrict__ out) {
  __shared__ double red[8][33];
  const int e = threadIdx.x & 31, q = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + e;
  double s = 0.0;
  if (i < nelem)
    for (int z = q; z < nslices; z += 8) s += part[(long long)z * nelem + i];
  red[q][e] = s;
  __syncthreads();
  if (q == 0 && i < nelem) {
    double v = red[0][e];
#pragma unroll
    for (int k = 1; k < 8; k++) v += red[k][e];
    out[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Kronecker core stage
// ---------------------------------------------------------------------------------------------
struct KronSizes { int n[8]; int off[8]; int nmodes; };

// prod_{m != skip} lambda_m[i_m] of the multi-index of `idx` (row-major), with the LAST mode's factor returned
// separately for idx and idx + 1 (pairs never straddle a last-mode row: callers use pairs only when n_last is even).
// 32-bit index arithmetic: a 64-bit div/mod per mode per element made the first version of these kernels ALU-bound
// (profiles/r01_kron_bench_v1.txt: kron_core at 11 % of HBM).
__device__ __forceinline__ double kron_prefix_prod(unsigned idx, const KronSizes& s, const double* __restrict__ lam, int skip,
                                                   unsigned& i_last) {
  const int last = s.nmodes - 1;
  i_last = idx % (unsigned)s.n[last];
  idx /= (unsigned)s.n[last];
  double p = 1.0;
#pragma unroll
  for (int m = 6; m >= 0; m--) {
    if (m < last) {
      const unsigned im = idx % (unsigned)s.n[m];
      idx /= (unsigned)s.n[m];
      if (m != skip) p *= lam[s.off[m] + im];
    }
  }
  return p;
}

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

// One pass: A = kron(lambda) + tau, h = T1 / A, and the four sums (sum log A, sum T1 h, sum 1/A, sum h^2).
// PAIRS: each thread handles element pairs with 16-byte loads/stores (total even, last mode even, total < 2^31).
template <bool PAIRS>
__global__ void __launch_bounds__(256) kron_core_kernel(const double* __restrict__ T1, const double* __restrict__ lam,
                                                        KronSizes s, const double* __restrict__ noise_inv, double add,
                                                        long long total, double* __restrict__ core, double* __restrict__ Aout,
                                                        double* __restrict__ part) {
  __shared__ double red[4][8];
  const double tau = (noise_inv ? noise_inv[0] : 0.0) + add;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  auto one = [&](double A, double t1, double& h) {
    const double ia = 1.0 / A;
    h = t1 * ia;
    s0 += log(A);
    s1 = fma(t1, h, s1);
    s2 += ia;
    s3 = fma(h, h, s3);
  };
  if (PAIRS) {
    const int last = s.nmodes - 1;
    const unsigned npairs = (unsigned)(total >> 1);
    for (unsigned pi = blockIdx.x * 256u + threadIdx.x; pi < npairs; pi += gridDim.x * 256u) {
      unsigned il;
      const double pre = kron_prefix_prod(2u * pi, s, lam, -1, il);
      const double2 t1 = reinterpret_cast<const double2*>(T1)[pi];
      const double A0 = fma(pre, lam[s.off[last] + il], tau), A1 = fma(pre, lam[s.off[last] + il + 1], tau);
      double2 h;
      one(A0, t1.x, h.x);
      one(A1, t1.y, h.y);
      if (core) reinterpret_cast<double2*>(core)[pi] = h;
      if (Aout) reinterpret_cast<double2*>(Aout)[pi] = make_double2(A0, A1);
    }
  } else {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
      const double A = kron_lambda_prod(i, s, lam, -1) + tau;
      double h;
      one(A, T1[i], h);
      if (core) core[i] = h;
      if (Aout) Aout[i] = A;
    }
  }
  // fixed-order block reduction: warp shuffles, then 8 warp partials in index order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; red[3][warp] = s3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int k = 0; k < 8; k++) v += red[threadIdx.x][k];
    part[(long long)blockIdx.x * 4 + threadIdx.x] = v;
  }
}

__global__ void kron_sums_finish_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ out) {
  // 4 warps, one per sum: lane-strided partial sums, then a shuffle tree (fixed order for a given nblocks)
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (q >= 4) return;
  double v = 0.0;
  for (int b = lane; b < nblocks; b += 32) v += part[(long long)b * 4 + q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) out[q] = v;
}

template <bool PAIRS>
__global__ void __launch_bounds__(256) kron_scale_kernel(const double* __restrict__ in, const double* __restrict__ lam,
                                                         KronSizes s, int skip, int power_inv_A,
                                                         const double* __restrict__ noise_inv, double add, long long total,
                                                         double* __restrict__ out) {
  const double tau = (noise_inv ? noise_inv[0] : 0.0) + add;
  if (PAIRS) {
    const int last = s.nmodes - 1;
    const unsigned npairs = (unsigned)(total >> 1);
    for (unsigned pi = blockIdx.x * 256u + threadIdx.x; pi < npairs; pi += gridDim.x * 256u) {
      unsigned il;
      const double pre_all = kron_prefix_prod(2u * pi, s, lam, -1, il);
      const double l0 = lam[s.off[last] + il], l1 = lam[s.off[last] + il + 1];
      double pre = pre_all;
      if (skip >= 0 && skip < last) { unsigned dummy; pre = kron_prefix_prod(2u * pi, s, lam, skip, dummy); }
      double2 v = in ? reinterpret_cast<const double2*>(in)[pi] : make_double2(1.0, 1.0);
      v.x *= (skip == last) ? pre : pre * l0;
      v.y *= (skip == last) ? pre : pre * l1;
      if (power_inv_A) { v.x /= fma(pre_all, l0, tau); v.y /= fma(pre_all, l1, tau); }
      reinterpret_cast<double2*>(out)[pi] = v;
    }
  } else {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
      double v = (in ? in[i] : 1.0) * kron_lambda_prod(i, s, lam, skip);
      if (power_inv_A) v /= (kron_lambda_prod(i, s, lam, -1) + tau);
      out[i] = v;
    }
  }
}

// c_k[j] = sum over the other modes' indices of  P / (lambda_k[j] P + tau),  P = prod_{m != k} lambda_m[i_m]:
// the diagonal weights of the closed-form gradient w.r.t. mode k's kernel matrix (tensorly_compat._KronNLL.backward).
// It depends on the eigenvalues only, so it is a pure reduction - the first version materialised W_k = P / A as a full
// tensor (kron_scale) and contracted it with a tensor of ones (mode_gram): 3 passes over 33 MB per mode at the C4 size.
// grid (n_k, S): block (j, y) sums its slice of the other-index space in a fixed order; kron_ck_finish adds the S partials.
__global__ void __launch_bounds__(256) kron_ck_kernel(const double* __restrict__ lam, KronSizes other, int lam_k_off,
                                                      const double* __restrict__ noise_inv, double add, long long other_total,
                                                      double* __restrict__ part) {
  __shared__ double red[8];
  const double tau = (noise_inv ? noise_inv[0] : 0.0) + add;
  const int j = blockIdx.x, S = gridDim.y;
  const double lk = lam[lam_k_off + j];
  double acc = 0.0;
  if (other_total < (1LL << 31)) {                  // 32-bit index arithmetic (a 64-bit div / mod per mode per term is ~10x the rest)
    const unsigned tot = (unsigned)other_total;
    for (unsigned o = blockIdx.y * 256u + threadIdx.x; o < tot; o += (unsigned)S * 256u) {
      unsigned idx = o;
      double P = 1.0;
#pragma unroll
      for (int m = 7; m >= 0; m--) {
        if (m < other.nmodes) {
          const unsigned nm = (unsigned)other.n[m], im = idx % nm;
          idx /= nm;
          P *= lam[other.off[m] + im];
        }
      }
      acc += P / fma(lk, P, tau);
    }
  } else {
    for (long long o = (long long)blockIdx.y * 256 + threadIdx.x; o < other_total; o += (long long)S * 256) {
      const double P = kron_lambda_prod(o, other, lam, -1);
      acc += P / fma(lk, P, tau);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int k = 0; k < 8; k++) v += red[k];
    part[(long long)j * S + blockIdx.y] = v;
  }
}
__global__ void kron_ck_finish_kernel(const double* __restrict__ part, int n, int S, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double v = 0.0;
  for (int y = 0; y < S; y++) v += part[(long long)j * S + y];
  out[j] = v;
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
// Jacobi eigensolver for n <= 128 (every per-mode matrix of the reference's configs): a CLUSTER OF TWO CTAs per
// matrix.  A single CTA is bound by shared-memory / L2 bandwidth (profiles/r01_kron_bench_v2.txt: 8 us per
// round-robin step at n = 128 with A in shared memory and V^T in L2; the first version, with serialised L2 round
// trips, 13 us).  Here CTA 0 keeps A in its shared memory and CTA 1 keeps V^T in its own; per step
//   CTA 0: the m/2 rotations of the step -> written to BOTH CTAs' rotation buffers (DSMEM stores), cluster barrier,
//          then ONE fused pass over A: every 2x2 block (pair k1 x pair k2) is read once, rotated from the left and
//          the right, and written once (half the shared-memory traffic of a row pass + a column pass);
//   CTA 1: cluster barrier, then the row rotations of V^T - concurrently with CTA 0's pass over A.
// Pairs follow the round-robin tournament in (a, b) order (not sorted), so consecutive pair slots touch consecutive
// rows/columns: conflict-free shared-memory access.  Rotation buffers are double-buffered by step parity; the one
// cluster barrier per step is the only cross-CTA synchronisation.
// ---------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 512;

__device__ __forceinline__ void rr_pair_ab(int step, int k, int m, int& a, int& b) {
  const int mm = m - 1;
  if (k == 0) { a = step % mm; b = mm; }
  else { a = (step + k) % mm; b = (step - k + mm) % mm; }
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// address of the same shared-memory variable in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t dsmem_addr(const void* local, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void dsmem_st_f64(uint32_t addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st_s32(uint32_t addr, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SC_THREADS, 1)
syevj_cluster_kernel(const double* __restrict__ Ain, int n, double* __restrict__ w, double* __restrict__ V,
                     int* __restrict__ info, int max_sweeps) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x >> 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_rank();
  const int m = (n + 1) & ~1, half = m / 2, lda = m + 1;
  double* mat = sm;                                   // rank 0: A [m][lda];  rank 1: V^T [m][lda] (rows = eigenvectors)
  double* rc = mat + (size_t)m * lda;                 // [2][half] cos, by step parity
  double* rs = rc + 2 * half;                         // [2][half] sin
  int* ra = reinterpret_cast<int*>(rs + 2 * half);    // [2][half] first index of the pair
  int* rb = ra + 2 * half;                            // [2][half] second index
  int* perm = rb + 2 * half;                          // [m] rank of eigenvalue i (ascending)
  int* flag = perm + m;                               // [2]: converged, pad
  double* red = reinterpret_cast<double*>(flag + 2);  // [SC_THREADS / 32]
  const double* A0 = Ain + (long long)b * n * n;

  double nrm_local = 0.0;
  for (int e = tid; e < m * m; e += SC_THREADS) {
    const int i = e / m, j = e - i * m;
    double v = 0.0;
    if (rank == 0) {
      // symmetrise from the UPPER triangle (torch.linalg.eigh(K, UPLO='U'), reference hogp.py:20)
      if (i < n && j < n) v = (j >= i) ? A0[(long long)i * n + j] : A0[(long long)j * n + i];
      nrm_local = fma(v, v, nrm_local);
    } else {
      v = (i == j) ? 1.0 : 0.0;
    }
    mat[i * lda + j] = v;
  }
  if (tid == 0) flag[0] = 0;
  auto block_sum = [&](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < SC_THREADS / 32; k++) t += red[k];
    return t;
  };
  const double norm2 = block_sum(nrm_local);          // meaningful in rank 0 only
  const uint32_t r_rc = dsmem_addr(rc, 1), r_rs = dsmem_addr(rs, 1), r_ra = dsmem_addr(ra, 1), r_rb = dsmem_addr(rb, 1);
  const uint32_t r_flag = dsmem_addr(flag, 1), r_perm = dsmem_addr(perm, 1);
  cluster_sync_all();                                  // both CTAs resident and initialised before any DSMEM store

  // loop-invariant thread -> work maps (see the update loops)
  const int k2u = tid % half, k1s = (tid < (SC_THREADS / half) * half) ? tid / half : half, k1step = SC_THREADS / half;
  const int vj = tid % n, vks = (tid < (SC_THREADS / n) * n) ? tid / n : half, vkstep = SC_THREADS / n;
  int converged = 0;
  if (rank == 0 && norm2 == 0.0) converged = 1;
  // a zero matrix converges immediately; publish that like any sweep result so that both CTAs agree
  if (rank == 0 && tid == 0 && converged) { flag[0] = 1; dsmem_st_s32(r_flag, 1); }
  cluster_sync_all();
  converged = flag[0];
  for (int sweep = 0; sweep < max_sweeps && !converged; sweep++) {
    double off_local = 0.0;
    for (int step = 0; step < m - 1; step++) {
      const int bo = (step & 1) * half;
      if (rank == 0 && tid < half) {
        int ia, ib; rr_pair_ab(step, tid, m, ia, ib);
        double c = 1.0, s = 0.0;
        if (ia < n && ib < n) {
          const double apq = mat[ia * lda + ib], app = mat[ia * lda + ia], aqq = mat[ib * lda + ib];
          off_local = fma(apq, apq, off_local);
          if (fabs(apq) > 1e-300 && fabs(apq) > 1e-19 * (fabs(app) + fabs(aqq))) {
            const double tau = (aqq - app) / (2.0 * apq);
            const double tt = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = rsqrt(1.0 + tt * tt);
            s = tt * c;
          }
        }
        const int o = bo + tid;
        rc[o] = c; rs[o] = s; ra[o] = ia; rb[o] = ib;
        dsmem_st_f64(r_rc + o * 8, c); dsmem_st_f64(r_rs + o * 8, s);
        dsmem_st_s32(r_ra + o * 4, ia); dsmem_st_s32(r_rb + o * 4, ib);
      }
      cluster_sync_all();
      if (rank == 0) {
        // fused two-sided update: block (k1, k2) = rows (a1, b1) x columns (a2, b2):  X' = G1 X G2^T, G = [[c, -s], [s, c]]
        // thread = (column pair k2 fixed, row pairs k1s, k1s + k1step, ..): no index division in the loop and the column
        // pair's rotation is loaded once per step
        if (k1s < half) {
         const double s2 = rs[bo + k2u], c2 = rc[bo + k2u];
         const int ca = ra[bo + k2u], cb = rb[bo + k2u];
         for (int k1 = k1s; k1 < half; k1 += k1step) {
          const double s1 = rs[bo + k1];
          if (s1 != 0.0 || s2 != 0.0) {
            const double c1 = rc[bo + k1];
            double* r0 = mat + ra[bo + k1] * lda;
            double* r1 = mat + rb[bo + k1] * lda;
            const double x00 = r0[ca], x01 = r0[cb], x10 = r1[ca], x11 = r1[cb];
            const double y00 = c1 * x00 - s1 * x10, y01 = c1 * x01 - s1 * x11;
            const double y10 = s1 * x00 + c1 * x10, y11 = s1 * x01 + c1 * x11;
            r0[ca] = c2 * y00 - s2 * y01; r0[cb] = s2 * y00 + c2 * y01;
            r1[ca] = c2 * y10 - s2 * y11; r1[cb] = s2 * y10 + c2 * y11;
          }
         }
        }
        __syncthreads();
      } else {
        for (int k = vks; k < half; k += vkstep) {
          const int j = vj;
          const double s = rs[bo + k];
          if (s != 0.0) {
            const double c = rc[bo + k];
            double* vp_ = mat + ra[bo + k] * lda + j;
            double* vq_ = mat + rb[bo + k] * lda + j;
            const double vp = *vp_, vq = *vq_;
            *vp_ = c * vp - s * vq;
            *vq_ = s * vp + c * vq;
          }
        }
      }
    }
    if (rank == 0) {
      const double off2 = block_sum(off_local);
      if (tid == 0) {
        const int cv = (off2 <= 1e-26 * norm2) ? 1 : 0;
        flag[0] = cv;
        dsmem_st_s32(r_flag, cv);
      }
    }
    cluster_sync_all();
    converged = flag[0];
  }
  if (rank == 0) {
    if (tid == 0 && !converged) info[b] = 1;
    // ascending order by rank, ties broken by index
    for (int i = tid; i < n; i += SC_THREADS) {
      const double li = mat[i * lda + i];
      int rk = 0;
      for (int j = 0; j < n; j++) {
        const double lj = mat[j * lda + j];
        rk += (lj < li) || (lj == li && j < i);
      }
      w[(long long)b * n + rk] = li;
      dsmem_st_s32(r_perm + i * 4, rk);
    }
  }
  cluster_sync_all();
  if (rank == 1) {
    for (int e = tid; e < n * n; e += SC_THREADS) {
      const int r = e / n, i = e - r * n;        // V[r][rank(i)] = V^T[i][r]
      V[(long long)b * n * n + (long long)r * n + perm[i]] = mat[i * lda + r];
    }
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

static bool kron_pairs_ok(const int* sizes, int nmodes, long long total, const void* a, const void* b, const void* c) {
  return nmodes >= 1 && nmodes <= 8 && !(sizes[nmodes - 1] & 1) && total < (1LL << 31) &&
         !(((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15);
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

static int ms_num_sms() {
  static PerDeviceInt per_dev;
  int& v = *per_dev.slot();
  if (!v) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

int ffgp_mode_dot_f64(const double* t, const double* mat, double* out, long long outer, int I, long long inner, int J,
                      int transpose_mat, void* stream) {
  if (!t || !mat || !out) return fail(-1, "ffgp_mode_dot_f64: null pointer%s", "");
  if (outer <= 0 || I <= 0 || inner <= 0 || J <= 0) return fail(-2, "ffgp_mode_dot_f64: bad size%s", "");
  const long long ncols = outer * inner;
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned16 = !(((uintptr_t)t | (uintptr_t)out) & 15);
  // (1) small factor matrix: streaming kernel, factor resident in shared memory
  if (I <= 128 && J <= 128 && aligned16 && ((inner == 1 && !(I & 1)) || (inner > 1 && !(inner & 1)))) {
    ModeSmallParams p;
    p.t = t; p.mat = mat; p.out = out; p.ncols = ncols; p.inner = inner; p.I = I; p.J = J; p.transpose_mat = transpose_mat;
    p.Ip = (I + 3) & ~3; p.Jp = (J + 7) & ~7;
    p.SA = p.Ip + ((4 - p.Ip % 16) + 16) % 16;             // smallest stride >= Ip with SA = 4 (mod 16)
    const int sms = ms_num_sms();
    const bool last = inner == 1;
    cudaError_t e;
    if (p.Jp <= 8) e = last ? launch_mode_small<true, 1>(p, sms, st) : launch_mode_small<false, 1>(p, sms, st);
    else if (p.Jp <= 16) e = last ? launch_mode_small<true, 2>(p, sms, st) : launch_mode_small<false, 2>(p, sms, st);
    else if (p.Jp <= 32) e = last ? launch_mode_small<true, 4>(p, sms, st) : launch_mode_small<false, 4>(p, sms, st);
    else e = last ? launch_mode_small<true, 8>(p, sms, st) : launch_mode_small<false, 8>(p, sms, st);
    FFGP_CUDA(e);
    ++ffgp::g_launches;
    FFGP_CUDA(cudaGetLastError());
    return 0;
  }
  // (2) large factor matrix (e.g. Tensor_linear 1024 -> 4096, gp_computation_pack.py:155-158): this IS a GEMM - the
  //     tiled DMMA GEMM of the dense path, batched over `outer`
  if ((I > 128 || J > 128) && I % 16 == 0 && J % 64 == 0) {
    if (inner == 1 && ncols % 64 == 0)          // out[c][j] = sum_k t[c][k] mat(j,k)
      return ffgp_gemm_f64(1, transpose_mat ? 0 : 1, t, I, 0, mat, transpose_mat ? J : I, 0, out, J, 0, (int)ncols, J, I, 1.0,
                           0.0, 0, 0, 1, stream);
    if (inner > 1 && inner % 64 == 0 && outer < (1 << 16))   // out_o[j][c] = sum_k mat(j,k) t_o[k][c]
      return ffgp_gemm_f64(transpose_mat ? 0 : 1, 0, mat, transpose_mat ? J : I, 0, t, (int)inner, (long long)I * inner, out,
                           (int)inner, (long long)J * inner, J, (int)inner, I, 1.0, 0.0, 0, 0, (int)outer, stream);
  }
  // (3) anything else: generic kernel
  const long long nchunks = (ncols + MD_COLS - 1) / MD_COLS;
  const int gx = (int)std::min<long long>(nchunks, 148LL * 8);
  mode_dot_kernel<<<dim3(gx, (J + 127) / 128), 256, 0, st>>>(t, mat, out, ncols, I, inner, J, transpose_mat);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  return 0;
}

static bool gram_stream_ok(const void* X, const void* Y, long long inner, int Ja, int Jb) {
  if (((uintptr_t)X | (uintptr_t)Y) & 15) return false;
  return inner == 1 ? (!(Ja & 1) && !(Jb & 1)) : !(inner & 1);
}

static int gram_slices(long long ncols, int Ja, int Jb, bool stream) {
  if (stream) {
    const long long tiles = (long long)((Ja + 31) / 32) * ((Jb + 31) / 32);
    const long long nchunks = (ncols + 63) / 64;                 // 8 warps x 8-column blocks per CTA pass
    const long long want = std::max<long long>(1, ((long long)ms_num_sms() * 2) / tiles);
    return (int)std::max<long long>(1, std::min(want, nchunks));
  }
  const long long tiles = (long long)((Ja + MG_T - 1) / MG_T) * ((Jb + MG_T - 1) / MG_T);
  long long want = std::max<long long>(1, (148LL * 4) / tiles);
  long long maxs = std::max<long long>(1, ncols / 256);
  return (int)std::min<long long>(std::min(want, maxs), 1024);
}

size_t ffgp_mode_gram_scratch_bytes(long long outer, long long inner, int Ja, int Jb) {
  // upper bound over both kernels (the streaming path is chosen by pointer alignment at call time)
  const int ns = std::max(gram_slices(outer * inner, Ja, Jb, true), gram_slices(outer * inner, Ja, Jb, false));
  return (size_t)ns * Ja * Jb * sizeof(double) + 256;
}

int ffgp_mode_gram_f64(const double* X, const double* Y, double* G, long long outer, long long inner, int Ja, int Jb,
                       void* scratch, size_t scratch_bytes, void* stream) {
  if (!X || !Y || !G || !scratch) return fail(-1, "ffgp_mode_gram_f64: null pointer%s", "");
  if (outer <= 0 || inner <= 0 || Ja <= 0 || Jb <= 0) return fail(-2, "ffgp_mode_gram_f64: bad size%s", "");
  if (scratch_bytes < ffgp_mode_gram_scratch_bytes(outer, inner, Ja, Jb)) return fail(-3, "ffgp_mode_gram_f64: scratch too small%s", "");
  const long long ncols = outer * inner;
  cudaStream_t st = (cudaStream_t)stream;
  const bool stream_path = gram_stream_ok(X, Y, inner, Ja, Jb);
  const int ns = gram_slices(ncols, Ja, Jb, stream_path);
  if (stream_path) {
    const int jm = std::min(32, std::max(Ja, Jb));
    double* pt = (double*)scratch;
    cudaError_t e;
    if (inner == 1) e = jm <= 8 ? launch_gram_stream<true, 1>(X, Y, pt, ncols, inner, Ja, Jb, ns, st)
                      : jm <= 16 ? launch_gram_stream<true, 2>(X, Y, pt, ncols, inner, Ja, Jb, ns, st)
                                 : launch_gram_stream<true, 4>(X, Y, pt, ncols, inner, Ja, Jb, ns, st);
    else e = jm <= 8 ? launch_gram_stream<false, 1>(X, Y, pt, ncols, inner, Ja, Jb, ns, st)
             : jm <= 16 ? launch_gram_stream<false, 2>(X, Y, pt, ncols, inner, Ja, Jb, ns, st)
                        : launch_gram_stream<false, 4>(X, Y, pt, ncols, inner, Ja, Jb, ns, st);
    FFGP_CUDA(e);
  } else {
    long long cps = (ncols + ns - 1) / ns;
    cps = (cps + MG_KC - 1) / MG_KC * MG_KC;
    mode_gram_kernel<<<dim3((Jb + MG_T - 1) / MG_T, (Ja + MG_T - 1) / MG_T, ns), 256, 0, st>>>(X, Y, (double*)scratch, ncols, inner,
                                                                                               Ja, Jb, cps);
    FFGP_CUDA(cudaGetLastError());
  }
  ++ffgp::g_launches;
  const long long nelem = (long long)Ja * Jb;
  gram_reduce_kernel<<<(unsigned)((nelem + 31) / 32), 256, 0, st>>>((const double*)scratch, ns, nelem, G);
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
  if (kron_pairs_ok(sizes_host, nmodes, total, T1, out_core, out_A))
    kron_core_kernel<true><<<nb, 256, 0, st>>>(T1, lambdas, s, noise_inv, add_scalar, total, out_core, out_A, (double*)scratch);
  else
    kron_core_kernel<false><<<nb, 256, 0, st>>>(T1, lambdas, s, noise_inv, add_scalar, total, out_core, out_A, (double*)scratch);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  kron_sums_finish_kernel<<<1, 128, 0, st>>>((const double*)scratch, nb, out_sums);
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
  if (kron_pairs_ok(sizes_host, nmodes, total, in, out, nullptr))
    kron_scale_kernel<true><<<nb, 256, 0, (cudaStream_t)stream>>>(in, lambdas, s, skip_mode, divide_by_A, noise_inv, add_scalar, total, out);
  else
    kron_scale_kernel<false><<<nb, 256, 0, (cudaStream_t)stream>>>(in, lambdas, s, skip_mode, divide_by_A, noise_inv, add_scalar, total, out);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  return 0;
}

size_t ffgp_kron_ck_scratch_bytes(int n_k) { return n_k > 0 ? (size_t)n_k * 64 * sizeof(double) + 256 : 0; }

int ffgp_kron_ck_f64(const double* lambdas, const int* sizes_host, int nmodes, int mode, const double* noise_inv,
                     double add_scalar, double* out, void* scratch, size_t scratch_bytes, void* stream) {
  if (!lambdas || !sizes_host || !out || !scratch) return fail(-1, "ffgp_kron_ck_f64: null pointer%s", "");
  if (nmodes < 2 || nmodes > 8 || mode < 0 || mode >= nmodes) return fail(-2, "ffgp_kron_ck_f64: 2..8 modes, 0 <= mode < nmodes%s", "");
  const int nk = sizes_host[mode];
  if (nk <= 0 || scratch_bytes < ffgp_kron_ck_scratch_bytes(nk)) return fail(-3, "ffgp_kron_ck_f64: scratch too small%s", "");
  KronSizes all = make_sizes(sizes_host, nmodes), other;
  memset(&other, 0, sizeof(other));
  long long other_total = 1;
  for (int m = 0; m < nmodes; m++) {
    if (m == mode) continue;
    other.n[other.nmodes] = all.n[m];
    other.off[other.nmodes] = all.off[m];
    other.nmodes++;
    other_total *= all.n[m];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int S = (int)std::max<long long>(1, std::min<long long>(64, (other_total + 2047) / 2048));
  kron_ck_kernel<<<dim3(nk, S), 256, 0, st>>>(lambdas, other, all.off[mode], noise_inv, add_scalar, other_total, (double*)scratch);
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  kron_ck_finish_kernel<<<(nk + 127) / 128, 128, 0, st>>>((const double*)scratch, nk, S, out);
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
  static PerDeviceOnce once;
  bool& attr = *once.slot();
  if (!attr) {
    FFGP_CUDA(cudaFuncSetAttribute(syevj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    FFGP_CUDA(cudaFuncSetAttribute(syevj_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  double* vt = (double*)workspace;
  FFGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * batch, st));
  // n <= 128: one-sided Jacobi with register-resident column pairs (syevj_hestenes.cuh).  FFGP_EIGH=twosided keeps the
  // round-1 two-sided cluster kernel (A/B comparisons only).
  static int hestenes = -1;
  if (hestenes < 0) { const char* e = getenv("FFGP_EIGH"); hestenes = (e && strcmp(e, "twosided") == 0) ? 0 : 1; }
  if (n <= HJ_MAX_N && hestenes) {
    static int sweeps = -1;              // FFGP_EIGH_SWEEPS=k: cap the sweeps (timing one sweep; results then unconverged)
    if (sweeps < 0) { const char* e = getenv("FFGP_EIGH_SWEEPS"); sweeps = e ? atoi(e) : 30; }
    FFGP_CUDA(hj_launch(A, n, batch, w, V, info, sweeps, st));
    ++ffgp::g_launches;
    return 0;
  }
  if (n <= SYEVJ_SMEM_N) {
    const int half = m / 2;
    const size_t smem = ((size_t)m * (m + 1) + 4 * half) * sizeof(double) + (size_t)(4 * half + m + 2) * sizeof(int) +
                        (SC_THREADS / 32) * sizeof(double) + 16;
    syevj_cluster_kernel<<<2 * batch, SC_THREADS, smem, st>>>(A, n, w, V, info, 40);
  } else {
    const size_t smem = (size_t)(m + SYEVJ_THREADS) * sizeof(double);
    double* wa = vt + (size_t)batch * n * n;
    syevj_kernel<<<batch, SYEVJ_THREADS, smem, st>>>(A, n, w, V, wa, vt, info, 40);
  }
  ++ffgp::g_launches;
  FFGP_CUDA(cudaGetLastError());
  return 0;
}

// Debug only (host sync): Jacobi sweeps of the most recent ffgp_syevj_f64 solve with n <= 128 (problem 0 of the batch).
int ffgp_debug_last_eigh_sweeps(void) {
  int v = -1;
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(&v, g_hj_last_sweeps, sizeof(int)) != cudaSuccess) return -1;
  return v;
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
  static PerDeviceOnce once;
  bool& attr = *once.slot();
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
