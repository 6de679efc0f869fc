// FP64 tensor-pipe GEMM for sm_100a (DMMA.8x8x4 through mma.sync.m8n8k4.f64).
//
// One kernel template serves every level-3 step of the dense GP path: the TRSM-as-GEMM
// against an explicit triangular inverse, the SYRK trailing update, the two triangular
// products of the recursive triangular inverse and the LAUUM-shaped M^T M product.  What
// varies is (a) which operand index is contiguous in memory, (b) the K range of a tile,
// which depends on the tile's row/column when one operand is triangular, and (c) whether
// only the lower-triangular tiles of C are produced.
//
// B200 notes (profiles/r01_fp64_peak_microbench.txt): every f64 mma shape lowers to
// DMMA.8x8x4; the pipe peaks at 37.1 TFLOP/s with >= 8 warps/SM and issues one DMMA per
// 16 clk per SM sub-partition, so per 128x128x16 k-block a CTA has ~4096 clk of math in
// which the next stages stream in through cp.async.  FP64 has no tcgen05/TMEM path.
//
// Layout: everything is row-major with leading dimension ld (elements).  All extents must
// be multiples of the tile (the host pads problems to a multiple of 128 with an identity
// tail, see dense_gp.cu), so there is no edge predication anywhere in the hot loop.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace ffgp {

// One-shot per-DEVICE initialisation flag.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to
// the current device's context: a process that drives a second GPU must set them again there.
struct PerDeviceOnce {
  bool done[64] = {};
  bool* slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return &done[dev & 63];
  }
};
struct PerDeviceInt {
  int v[64] = {};
  int* slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return &v[dev & 63];
  }
};

enum KMode : int {
  K_FULL = 0,       // k in [0, K)
  K_LE_ROW = 1,     // A lower-triangular in (i,p): k in [0, i0+BM)
  K_LE_COL = 2,     // B given as B[j][p] lower-triangular (p <= j): k in [0, j0+BN)
  K_GE_COL = 3,     // B given as B[p][j] lower-triangular (p >= j): k in [j0, K)
  K_GE_ROW = 4      // A given as A[p][i] lower-triangular (p >= i): k in [i0, K)
};

struct GemmParams {
  const double* A;
  const double* B;
  double* C;
  int M, N, K;
  int lda, ldb, ldc;
  long long sA, sB, sC;   // batch strides in elements
  int inner;              // blockIdx.z = outer * inner + j: a second (inner) batch level, e.g. the nodes of one
  long long iA, iB, iC;   //   level of the bottom-up triangular inverse inside each problem of the outer batch
  double alpha, beta;     // C = alpha * A.B + beta * C
  int lower_only;         // 1: only tiles with tj <= ti (needs BM == BN, M == N)
  int kmode;
  int heavy_first;        // 1: launch the longest-K tiles first
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// BM x BN C tile, BK k-block, WARPS_M x WARPS_N warps, STAGES-deep cp.async ring.
// A_KMAJ: A(i,p) = A[i*lda + p]   else  A(i,p) = A[p*lda + i]
// B_KMAJ: B(p,j) = B[j*ldb + p]   else  B(p,j) = B[p*ldb + j]
template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES, bool A_KMAJ, bool B_KMAJ>
struct GemmCfg {
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  static constexpr int MT = WM / 8, NT = WN / 8;
  // padded smem strides (doubles): k-major rows are BK+4 so that the 4 rows x 4 k of a
  // half-warp fragment load cover all 32 banks once; m-major rows are BM+4 for the same reason.
  static constexpr int A_ROWS = A_KMAJ ? BM : BK;
  static constexpr int A_LD = A_KMAJ ? (BK + 4) : (BM + 4);
  static constexpr int B_ROWS = B_KMAJ ? BN : BK;
  static constexpr int B_LD = B_KMAJ ? (BK + 4) : (BN + 4);
  static constexpr int A_STAGE = A_ROWS * A_LD;
  static constexpr int B_STAGE = B_ROWS * B_LD;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
};

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES, bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32)
gemm_dmma_kernel(const GemmParams p) {
  using Cfg = GemmCfg<BM, BN, BK, WARPS_M, WARPS_N, STAGES, A_KMAJ, B_KMAJ>;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * Cfg::A_STAGE;

  // ---- tile coordinates -------------------------------------------------------------
  const int tiles_m = p.M / BM, tiles_n = p.N / BN;
  int t = blockIdx.x;
  int ti, tj;
  if (p.lower_only) {
    const int total = tiles_m * (tiles_m + 1) / 2;
    // heavy_first with K_GE_ROW wants small ti first (natural order); with K_LE_ROW /
    // K_FULL-on-lower the later rows are the long ones, so walk backwards.
    if (p.heavy_first && p.kmode != K_GE_ROW) t = total - 1 - t;
    ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (ti * (ti + 1) / 2 > t) --ti;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    tj = t - ti * (ti + 1) / 2;
  } else {
    const int total = tiles_m * tiles_n;
    if (p.heavy_first && (p.kmode == K_LE_ROW || p.kmode == K_LE_COL)) t = total - 1 - t;
    if (p.kmode == K_LE_COL || p.kmode == K_GE_COL) {   // K length depends on the column
      tj = t / tiles_m; ti = t - tj * tiles_m;
    } else {
      ti = t / tiles_n; tj = t - ti * tiles_n;
    }
  }
  const int i0 = ti * BM, j0 = tj * BN;
  int k_lo = 0, k_hi = p.K;
  if (p.kmode == K_LE_ROW) k_hi = min(p.K, i0 + BM);
  else if (p.kmode == K_LE_COL) k_hi = min(p.K, j0 + BN);
  else if (p.kmode == K_GE_COL) k_lo = j0;
  else if (p.kmode == K_GE_ROW) k_lo = i0;
  const int KT = (k_hi - k_lo) / BK;

  const int zo = blockIdx.z / p.inner, zi = blockIdx.z - zo * p.inner;
  const double* __restrict__ Ag = p.A + (long long)zo * p.sA + (long long)zi * p.iA;
  const double* __restrict__ Bg = p.B + (long long)zo * p.sB + (long long)zi * p.iB;
  double* __restrict__ Cg = p.C + (long long)zo * p.sC + (long long)zi * p.iC;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int g = lane >> 2, tq = lane & 3;

  // ---- global -> shared stage loader (16-byte cp.async, fully coalesced) ---------------
  auto load_stage = [&](int stage, int kt) {
    const int k0 = k_lo + kt * BK;
    double* as = As + stage * Cfg::A_STAGE;
    double* bs = Bs + stage * Cfg::B_STAGE;
    if (A_KMAJ) {
      constexpr int CH = BK / 2;                         // 16B chunks per row
      for (int c = tid; c < BM * CH; c += Cfg::THREADS) {
        const int r = c / CH, q = c - r * CH;
        cp_async16(as + r * Cfg::A_LD + q * 2, Ag + (long long)(i0 + r) * p.lda + k0 + q * 2);
      }
    } else {
      constexpr int CH = BM / 2;
      for (int c = tid; c < BK * CH; c += Cfg::THREADS) {
        const int r = c / CH, q = c - r * CH;
        cp_async16(as + r * Cfg::A_LD + q * 2, Ag + (long long)(k0 + r) * p.lda + i0 + q * 2);
      }
    }
    if (B_KMAJ) {
      constexpr int CH = BK / 2;
      for (int c = tid; c < BN * CH; c += Cfg::THREADS) {
        const int r = c / CH, q = c - r * CH;
        cp_async16(bs + r * Cfg::B_LD + q * 2, Bg + (long long)(j0 + r) * p.ldb + k0 + q * 2);
      }
    } else {
      constexpr int CH = BN / 2;
      for (int c = tid; c < BK * CH; c += Cfg::THREADS) {
        const int r = c / CH, q = c - r * CH;
        cp_async16(bs + r * Cfg::B_LD + q * 2, Bg + (long long)(k0 + r) * p.ldb + j0 + q * 2);
      }
    }
  };

  double acc[Cfg::MT][Cfg::NT][2];
#pragma unroll
  for (int i = 0; i < Cfg::MT; i++)
#pragma unroll
    for (int j = 0; j < Cfg::NT; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < KT) load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const double* as = As + (kt % STAGES) * Cfg::A_STAGE;
    const double* bs = Bs + (kt % STAGES) * Cfg::B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK / 4; kk++) {
      double af[Cfg::MT], bf[Cfg::NT];
#pragma unroll
      for (int i = 0; i < Cfg::MT; i++) {
        const int r = wm * Cfg::WM + i * 8 + g;
        af[i] = A_KMAJ ? as[r * Cfg::A_LD + kk * 4 + tq] : as[(kk * 4 + tq) * Cfg::A_LD + r];
      }
#pragma unroll
      for (int j = 0; j < Cfg::NT; j++) {
        const int c = wn * Cfg::WN + j * 8 + g;
        bf[j] = B_KMAJ ? bs[c * Cfg::B_LD + kk * 4 + tq] : bs[(kk * 4 + tq) * Cfg::B_LD + c];
      }
#pragma unroll
      for (int i = 0; i < Cfg::MT; i++)
#pragma unroll
        for (int j = 0; j < Cfg::NT; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: C fragment (row g, cols 2*tq, 2*tq+1) -> 16-byte stores ----------------
  const double alpha = p.alpha, beta = p.beta;
#pragma unroll
  for (int i = 0; i < Cfg::MT; i++) {
    const int r = i0 + wm * Cfg::WM + i * 8 + g;
#pragma unroll
    for (int j = 0; j < Cfg::NT; j++) {
      const int c = j0 + wn * Cfg::WN + j * 8 + tq * 2;
      double2* dst = reinterpret_cast<double2*>(Cg + (long long)r * p.ldc + c);
      double2 v;
      v.x = alpha * acc[i][j][0];
      v.y = alpha * acc[i][j][1];
      if (beta != 0.0) {
        const double2 o = *dst;
        v.x += beta * o.x;
        v.y += beta * o.y;
      }
      *dst = v;
    }
  }
}

// Host launcher.  big = 128x128 tiles (one CTA per SM); small = 64x64x16 tiles on 4 warps for the levels where
// 128-tiles would leave most of the 148 SMs idle.  The big-tile variant (warp grid, BK, stages) is selectable
// with FFGP_GEMM_CFG for tuning; the default is the one measured fastest (see profiles/).
template <int BM, int BN, int BK, int WM_, int WN_, int ST, bool A_KMAJ, bool B_KMAJ>
cudaError_t launch_cfg(const GemmParams& p, int batch, cudaStream_t st) {
  using Cfg = GemmCfg<BM, BN, BK, WM_, WN_, ST, A_KMAJ, B_KMAJ>;
  auto kern = gemm_dmma_kernel<BM, BN, BK, WM_, WN_, ST, A_KMAJ, B_KMAJ>;
  static PerDeviceOnce once;
  bool& attr_set = *once.slot();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tm = p.M / BM, tn = p.N / BN;
  const int tiles = p.lower_only ? tm * (tm + 1) / 2 : tm * tn;
  if (tiles == 0 || batch == 0) return cudaSuccess;
  kern<<<dim3(tiles, 1, batch), Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p);
  return cudaGetLastError();
}

inline int gemm_big_cfg() {
  static int cfg = -1;
  if (cfg < 0) {
    const char* e = getenv("FFGP_GEMM_CFG");
    cfg = e ? atoi(e) : 2;
    if (cfg < 0 || cfg > 3) cfg = 2;
  }
  return cfg;
}

template <bool A_KMAJ, bool B_KMAJ>
cudaError_t launch_gemm(const GemmParams& p, int batch, bool big, cudaStream_t st) {
  if (!big) return launch_cfg<64, 64, 16, 2, 2, 3, A_KMAJ, B_KMAJ>(p, batch, st);
  switch (gemm_big_cfg()) {
    case 1: return launch_cfg<128, 128, 16, 4, 4, 3, A_KMAJ, B_KMAJ>(p, batch, st);   // 16 warps, 32x32 warp tiles
    case 2: return launch_cfg<128, 128, 32, 2, 4, 2, A_KMAJ, B_KMAJ>(p, batch, st);   // BK 32, double buffer
    case 3: return launch_cfg<128, 128, 32, 4, 4, 2, A_KMAJ, B_KMAJ>(p, batch, st);
    default: return launch_cfg<128, 128, 16, 2, 4, 3, A_KMAJ, B_KMAJ>(p, batch, st);  // 8 warps, 64x32 warp tiles
  }
}

}  // namespace ffgp
