// One-launch potrf + trtri of a 256 x 256 diagonal block:  (L, M = L^-1) of A[off:off+256, off:off+256].
//
// Why (profiles/r01_timeline_c2_v2.txt): on the single large factorisation (BASELINE config 2, N = 8192) every
// right-looking step waits for the serial chain of its 256-block,
//     base(A11) ; L21 = A21 M11^T ; A22 -= L21 L21^T ; base(A22) ; T = M22 L21 ; M21 = -T M11,
// which as six dependent launches (two 48 us base kernels + four single-tile GEMM launches of 25-47 us each: launch
// latency, tensor-map fetch, pipeline fill and a 128 KB round trip through L2 per product) costs ~300 us against
// ~30 us of tensor-pipe work.  Here ONE CTA runs the whole chain and the operands never leave the SM:
//   * the 128-block being factored / its factor and inverse live in shared memory in the base kernel's layout
//     T[128][129] (base_factor_smem, dense_kernels.cuh), and the next product reads its fragments straight from there;
//   * the operand that must come from global memory (A21, and L21 / M11 re-read from L2 for the two products of the
//     inverse) streams through a 4-slot cp.async ring of 128 x 16 panels;
//   * products run on the FP64 tensor pipe (DMMA.8x8x4), 8 warps of 64 x 32 with the same cyclic fragment ownership
//     and compile-time live-fragment variants as gemm_tma_kernel, so the structural zeros of the three triangular
//     products and of the symmetric update are skipped.
// The batched path (BASELINE config 5, N = 512) uses the same kernel with one CTA per problem: 2 launches instead of 12
// for the two 256-blocks of a problem.
#pragma once
#include "dense_kernels.cuh"
#include "gemm_dmma.cuh"

namespace ffgp {

constexpr int F2_LD = BASE_N + 1;                  // row stride of the base layout T
constexpr int F2_XLD = 132;                        // row stride of a plain 128 x 128 operand parked in the T region
constexpr int F2_KLD = 20, F2_NLD = 132;           // ring slot strides: k-major rows [128][20], n-major rows [16][132]
constexpr int F2_SLOT = 128 * F2_KLD;              // doubles per ring slot (>= 16 * F2_NLD)
constexpr int F2_SLOTS = 4;
constexpr int F2_RING_OFF = (int)(base_smem_bytes(BASE_N) / sizeof(double));
constexpr size_t F2_SMEM = base_smem_bytes(BASE_N) + (size_t)F2_SLOTS * F2_SLOT * sizeof(double);
static_assert(128 * F2_XLD <= F2_RING_OFF, "parked operand must fit below the ring");
static_assert(F2_SMEM <= 232448, "shared memory budget of one CTA");

enum : int { F2_STREAM = 0, F2_PLAIN = 1, F2_INV = 2 };

// live fragments of a k16 step: i in [ILO, IHI), j in [JLO, JHI); ROLE >= 0: only fragments on/below the diagonal of
// the (symmetric) output tile for the warp wm = ROLE >> 2, wn = ROLE & 3
template <int ILO, int IHI, int JLO, int JHI, int ROLE>
struct F2Live {
  __device__ static constexpr bool live(int i, int j) {
    return i >= ILO && i < IHI && j >= JLO && j < JHI && (ROLE < 0 || 2 * i + (ROLE >> 2) - 4 * j - (ROLE & 3) >= 0);
  }
  __device__ static constexpr bool row_live(int i) { return live(i, 0) || live(i, 1) || live(i, 2) || live(i, 3); }
  __device__ static constexpr bool col_live(int j) {
    return live(0, j) || live(1, j) || live(2, j) || live(3, j) || live(4, j) || live(5, j) || live(6, j) || live(7, j);
  }
};

// One k16 step of a 128 x 128 product.  Operand sources:
//   A: F2_STREAM  slot[r][p - k0]  (k-major panel)      B: F2_STREAM  slot[p - k0][c]  (n-major panel)
//      F2_PLAIN   X[r][p]          (stride F2_XLD)          F2_PLAIN   X[c][p]          (B = X^T)
//      F2_INV     Minv[r][p], p <= r, from the base layout T[p][r + 1]
//                                                           F2_INV     Minv[c][p], p <= c  (B = Minv^T)
// Structural zeros: compile-time live-fragment sets (Lv), one straight-line body per set, selected per k16 step by a
// warp-uniform switch.  Run-time bounds around fragment rows / columns inside ONE body were tried first: ptxas turns the
// warp-uniform branches into predicated DMMAs, and a predicated-off DMMA keeps its issue slot - a triangular product
// then costs exactly what the dense one does (4500 clk per k16 step either way, tools/f256_trace.cu).
template <int ASRC, int BSRC, bool NEG_A, class Lv>
__device__ __forceinline__ void f2_kstep(double (&acc)[8][4][2], const double* __restrict__ T,
                                         const double* __restrict__ slot, const int k0, const int wm, const int wn,
                                         const int g, const int tq) {
  // fragments are double-buffered in registers one k4 step ahead: the loads of step kk + 1 are in flight behind the DMMAs
  // of step kk (a single CTA per SM has only two warps per sub-partition to hide a shared-memory round trip otherwise)
  double af[2][8], bf[2][4];
  auto load_frags = [&](const int buf, const int kk) {
    const int pl = kk * 4 + tq, p = k0 + pl;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (!Lv::row_live(i)) continue;
      const int r = (2 * i + wm) * 8 + g;
      if (ASRC == F2_STREAM) af[buf][i] = slot[r * F2_KLD + pl];
      else if (ASRC == F2_PLAIN) af[buf][i] = NEG_A ? -T[r * F2_XLD + p] : T[r * F2_XLD + p];
      else { const double v = T[p * F2_LD + r + 1]; af[buf][i] = (p <= r) ? v : 0.0; }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (!Lv::col_live(j)) continue;
      const int c = (4 * j + wn) * 8 + g;
      if (BSRC == F2_STREAM) bf[buf][j] = slot[pl * F2_NLD + c];
      else if (BSRC == F2_PLAIN) bf[buf][j] = T[c * F2_XLD + p];
      else { const double v = T[p * F2_LD + c + 1]; bf[buf][j] = (p <= c) ? v : 0.0; }
    }
  };
  load_frags(0, 0);
#pragma unroll
  for (int kk = 0; kk < 4; kk++) {
    const int cur = kk & 1;
    if (kk < 3) load_frags(cur ^ 1, kk + 1);
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (Lv::live(i, j)) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
  }
}

// 128 x 16 panel loads into a ring slot (16-byte cp.async, coalesced)
__device__ __forceinline__ void f2_load_k(double* slot, const double* __restrict__ X, int ldx, int k0, int tid) {
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int c = tid + it * 256, r = c >> 3, q = c & 7;
    cp_async16(slot + r * F2_KLD + q * 2, X + (long long)r * ldx + k0 + q * 2);
  }
}
__device__ __forceinline__ void f2_load_n(double* slot, const double* __restrict__ X, int ldx, int k0, int tid) {
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int c = tid + it * 256, r = c >> 6, q = c & 63;
    cp_async16(slot + r * F2_NLD + q * 2, X + (long long)(k0 + r) * ldx + q * 2);
  }
}

__device__ __forceinline__ void f2_zero(double (&acc)[8][4][2]) {
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}

// accumulators -> plain 128 x 128 operand X (stride F2_XLD) in the T region
__device__ __forceinline__ void f2_park(const double (&acc)[8][4][2], double* __restrict__ X, int wm, int wn, int g, int tq) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int r = (2 * i + wm) * 8 + g;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int c = (4 * j + wn) * 8 + 2 * tq;
      *reinterpret_cast<double2*>(X + r * F2_XLD + c) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

// base-layout block in T -> L and M (128 x 128, zeros above the diagonal) and the block's log-determinant slot
__device__ __forceinline__ void f2_store_base(const double* __restrict__ T, double* __restrict__ red, double* __restrict__ L,
                                              double* __restrict__ M, int ld, double* __restrict__ logdet_slot, int tid) {
  for (int e = tid; e < BASE_N * BASE_N; e += 256) {
    const int i = e / BASE_N, k = e % BASE_N;
    L[(long long)i * ld + k] = (k <= i) ? T[i * F2_LD + k] : 0.0;
    M[(long long)i * ld + k] = (k <= i) ? T[k * F2_LD + i + 1] : 0.0;
  }
  if (tid < 64) red[tid] = log(T[tid * F2_LD + tid]) + log(T[(tid + 64) * F2_LD + tid + 64]);
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int k = 0; k < 64; k++) s += red[k];
    *logdet_slot = s;
  }
}

// base_factor_smem is called twice per 256-block.  Inlined twice, ptxas schedules the two copies differently inside this
// (244-register) kernel and the second copy's update loop ran 40 % slower than the stand-alone base kernel
// (tools/f256_trace.cu: 89.7k clk against 68.7k for the first copy); one out-of-line copy serves both calls.
__device__ __noinline__ int f2_base_factor(double* __restrict__ sm, const int tid) { return base_factor_smem<BASE_N>(sm, tid); }

#ifdef FFGP_F256_TRACE
__device__ long long g_f256_trace[16];                  // clock64 at the phase boundaries (tools/f256_trace.cu)
#define F2_STAMP(k) do { __syncthreads(); if (threadIdx.x == 0) g_f256_trace[k] = clock64(); } while (0)
__device__ long long g_f256_trace2[8 * 8 * 3];          // P5: [warp][kt][before wait | after barrier | after the k-step]
#define F2_STAMP2(kt, s) do { if (lane == 0) g_f256_trace2[(warp * 8 + (kt)) * 3 + (s)] = clock64(); } while (0)
#else
#define F2_STAMP2(kt, s) do { } while (0)
#define F2_STAMP(k) do { } while (0)
#endif

// `skip_tri` = 0 computes every fragment of every product (reference behaviour for A/B tests, FFGP_F256=2)
__global__ void __launch_bounds__(256, 1)
factor256_kernel(const double* __restrict__ A, double* __restrict__ L, double* __restrict__ M, int ld, long long sbatch,
                 double* __restrict__ logdet_part, int logdet_stride, int blk, int* __restrict__ info, int row_offset,
                 int skip_tri) {
  extern __shared__ __align__(16) double sm[];
  double* T = sm;
  double* red = sm + BASE_N * F2_LD + 2 * BASE_F + 64;
  double* ring = sm + F2_RING_OFF;
  const int tid = threadIdx.x, b = blockIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = (warp & 3) ^ (wm ? 3 : 0);     // same warp -> fragment map as gemm_tma_kernel
  const int g = lane >> 2, tq = lane & 3, role = wm * 4 + wn;
  A += b * sbatch; L += b * sbatch; M += b * sbatch;
  const long long o21 = (long long)BASE_N * ld, o22 = o21 + BASE_N;
  const double* A21 = A + o21;
  double* logdet = logdet_part + (long long)b * logdet_stride + blk;
  double acc[8][4][2];

  F2_STAMP(0);
  // ---------------- P1: (L11, M11) = base(A11) ----------------
  for (int e = tid; e < BASE_N * BASE_N; e += 256) {
    const int i = e / BASE_N, k = e % BASE_N;
    if (k <= i) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&T[i * F2_LD + k]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(A + (long long)i * ld + k));
      T[k * F2_LD + i + 1] = (i == k) ? 1.0 : 0.0;
    }
  }
  cp_async_commit();
  // the first panels of A21 stream in behind the factorisation of A11
#pragma unroll
  for (int s = 0; s < F2_SLOTS - 1; s++) { f2_load_k(ring + s * F2_SLOT, A21, ld, 16 * s, tid); cp_async_commit(); }
  cp_async_wait<F2_SLOTS - 1>();
  __syncthreads();
  F2_STAMP(1);
  const int fail1 = f2_base_factor(sm, tid);
  __syncthreads();
  F2_STAMP(2);
  f2_store_base(T, red, L, M, ld, logdet, tid);
  F2_STAMP(3);

  // ---------------- P2: L21 = A21 M11^T   (A streamed k-major, B = M11 from T, p <= col) ----------------
  f2_zero(acc);
#pragma unroll 1
  for (int kt = 0; kt < 8; kt++) {
    cp_async_wait<F2_SLOTS - 2>();
    __syncthreads();
    if (kt + F2_SLOTS - 1 < 8) f2_load_k(ring + ((kt + F2_SLOTS - 1) % F2_SLOTS) * F2_SLOT, A21, ld, 16 * (kt + F2_SLOTS - 1), tid);
    cp_async_commit();
    const double* slot = ring + (kt % F2_SLOTS) * F2_SLOT;
    const int jlo = skip_tri ? (2 * kt - wn + 3) >> 2 : 0;       // fragment j is live iff 4 j + wn >= 2 kt
    switch (jlo) {
      case 0: f2_kstep<F2_STREAM, F2_INV, false, F2Live<0, 8, 0, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 1: f2_kstep<F2_STREAM, F2_INV, false, F2Live<0, 8, 1, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 2: f2_kstep<F2_STREAM, F2_INV, false, F2Live<0, 8, 2, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 3: f2_kstep<F2_STREAM, F2_INV, false, F2Live<0, 8, 3, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      default: break;
    }
  }
  cp_async_wait<0>();
  __syncthreads();                                   // every warp is done with M11 in T
  {
    double* L21 = L + o21;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int r = (2 * i + wm) * 8 + g;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = (4 * j + wn) * 8 + 2 * tq;
        *reinterpret_cast<double2*>(L21 + (long long)r * ld + c) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
    f2_park(acc, T, wm, wn, g, tq);
  }
  F2_STAMP(4);
  // ---------------- P3: A22' = A22 - L21 L21^T  (both operands parked in the T region; lower part only) ----------------
  {
    const double* A22 = A + o22;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int r = (2 * i + wm) * 8 + g;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = (4 * j + wn) * 8 + 2 * tq;
        double2 o = make_double2(0.0, 0.0);
        if (2 * i + wm - 4 * j - wn >= 0) o = *reinterpret_cast<const double2*>(A22 + (long long)r * ld + c);
        acc[i][j][0] = o.x; acc[i][j][1] = o.y;
      }
    }
  }
  __syncthreads();                                   // L21 parked
#pragma unroll 1
  for (int kt = 0; kt < 8; kt++) {
    switch (skip_tri ? role : -1) {
      case 0: f2_kstep<F2_PLAIN, F2_PLAIN, true, F2Live<0, 8, 0, 4, 0>>(acc, T, nullptr, 16 * kt, wm, wn, g, tq); break;
      case 1: f2_kstep<F2_PLAIN, F2_PLAIN, true, F2Live<0, 8, 0, 4, 1>>(acc, T, nullptr, 16 * kt, wm, wn, g, tq); break;
      case 2: f2_kstep<F2_PLAIN, F2_PLAIN, true, F2Live<0, 8, 0, 4, 2>>(acc, T, nullptr, 16 * kt, wm, wn, g, tq); break;
      case 3: f2_kstep<F2_PLAIN, F2_PLAIN, true, F2Live<0, 8, 0, 4, 3>>(acc, T, nullptr, 16 * kt, wm, wn, g, tq); break;
      case 4: f2_kstep<F2_PLAIN, F2_PLAIN, true, F2Live<0, 8, 0, 4, 4>>(acc, T, nullptr, 16 * kt, wm, wn, g, tq); break;
      case 5: f2_kstep<F2_PLAIN, F2_PLAIN, true, F2Live<0, 8, 0, 4, 5>>(acc, T, nullptr, 16 * kt, wm, wn, g, tq); break;
      case 6: f2_kstep<F2_PLAIN, F2_PLAIN, true, F2Live<0, 8, 0, 4, 6>>(acc, T, nullptr, 16 * kt, wm, wn, g, tq); break;
      case 7: f2_kstep<F2_PLAIN, F2_PLAIN, true, F2Live<0, 8, 0, 4, 7>>(acc, T, nullptr, 16 * kt, wm, wn, g, tq); break;
      default: f2_kstep<F2_PLAIN, F2_PLAIN, true, F2Live<0, 8, 0, 4, -1>>(acc, T, nullptr, 16 * kt, wm, wn, g, tq); break;
    }
  }
  __syncthreads();                                   // every warp is done with the parked L21
  // the updated block goes back into the base layout: lower part from the accumulators, inverse part = identity
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int r = (2 * i + wm) * 8 + g;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (2 * i + wm - 4 * j - wn < 0) continue;
      const int c = (4 * j + wn) * 8 + 2 * tq;
      if (c <= r) T[r * F2_LD + c] = acc[i][j][0];
      if (c + 1 <= r) T[r * F2_LD + c + 1] = acc[i][j][1];
    }
  }
  // the first panels of L21 (re-read from L2, n-major) stream in behind the factorisation of A22'
  const double* L21g = L + o21;
#pragma unroll
  for (int s = 0; s < F2_SLOTS - 1; s++) { f2_load_n(ring + s * F2_SLOT, L21g, ld, 16 * s, tid); cp_async_commit(); }
  __syncthreads();                                   // lower part complete before the identity goes in above it
  for (int e = tid; e < BASE_N * BASE_N; e += 256) {
    const int i = e / BASE_N, k = e % BASE_N;
    if (k <= i) T[k * F2_LD + i + 1] = (i == k) ? 1.0 : 0.0;
  }
  __syncthreads();
  F2_STAMP(5);
  const int fail2 = f2_base_factor(sm, tid);
  __syncthreads();
  F2_STAMP(6);
  f2_store_base(T, red, L + o22, M + o22, ld, logdet + BASE_N / BASE_N_BATCHED, tid);
  if (tid == 0) {
    const int fc = fail1 >= 0 ? fail1 : (fail2 >= 0 ? BASE_N + fail2 : -1);
    if (fc >= 0) atomicCAS(info + b, 0, row_offset + fc + 1);
  }

  F2_STAMP(7);
  // ---------------- P4: T' = M22 L21   (A = M22 from T, p <= row; B streamed n-major) ----------------
  f2_zero(acc);
#pragma unroll 1
  for (int kt = 0; kt < 8; kt++) {
    cp_async_wait<F2_SLOTS - 2>();
    __syncthreads();
    if (kt + F2_SLOTS - 1 < 8) f2_load_n(ring + ((kt + F2_SLOTS - 1) % F2_SLOTS) * F2_SLOT, L21g, ld, 16 * (kt + F2_SLOTS - 1), tid);
    cp_async_commit();
    const double* slot = ring + (kt % F2_SLOTS) * F2_SLOT;
    switch (skip_tri ? kt : 0) {                     // fragment i is live iff 2 i + wm >= 2 kt, i.e. i >= kt
      case 0: f2_kstep<F2_INV, F2_STREAM, false, F2Live<0, 8, 0, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 1: f2_kstep<F2_INV, F2_STREAM, false, F2Live<1, 8, 0, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 2: f2_kstep<F2_INV, F2_STREAM, false, F2Live<2, 8, 0, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 3: f2_kstep<F2_INV, F2_STREAM, false, F2Live<3, 8, 0, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 4: f2_kstep<F2_INV, F2_STREAM, false, F2Live<4, 8, 0, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 5: f2_kstep<F2_INV, F2_STREAM, false, F2Live<5, 8, 0, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 6: f2_kstep<F2_INV, F2_STREAM, false, F2Live<6, 8, 0, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      default: f2_kstep<F2_INV, F2_STREAM, false, F2Live<7, 8, 0, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
    }
  }
  cp_async_wait<0>();
  __syncthreads();                                   // every warp is done with M22 in T and with the ring
  // the first panels of M11 (re-read from L2, n-major) are requested before T' is parked
  const double* M11g = M;
#pragma unroll
  for (int s = 0; s < F2_SLOTS - 1; s++) { f2_load_n(ring + s * F2_SLOT, M11g, ld, 16 * s, tid); cp_async_commit(); }
  f2_park(acc, T, wm, wn, g, tq);

  F2_STAMP(8);
  // ---------------- P5: M21 = -T' M11   (A = T' parked; B streamed n-major, p >= col) ----------------
  f2_zero(acc);
#pragma unroll 1
  for (int kt = 0; kt < 8; kt++) {
    F2_STAMP2(kt, 0);
    cp_async_wait<F2_SLOTS - 2>();
    __syncthreads();                                 // (kt = 0: also publishes the parked T')
    F2_STAMP2(kt, 1);
    if (kt + F2_SLOTS - 1 < 8) f2_load_n(ring + ((kt + F2_SLOTS - 1) % F2_SLOTS) * F2_SLOT, M11g, ld, 16 * (kt + F2_SLOTS - 1), tid);
    cp_async_commit();
    const double* slot = ring + (kt % F2_SLOTS) * F2_SLOT;
    const int jhi = skip_tri ? ((2 * kt + 1 - wn + 4) >> 2) : 4;   // fragment j is live iff 4 j + wn <= 2 kt + 1
    switch (jhi) {
      case 1: f2_kstep<F2_PLAIN, F2_STREAM, false, F2Live<0, 8, 0, 1, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 2: f2_kstep<F2_PLAIN, F2_STREAM, false, F2Live<0, 8, 0, 2, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 3: f2_kstep<F2_PLAIN, F2_STREAM, false, F2Live<0, 8, 0, 3, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      case 4: f2_kstep<F2_PLAIN, F2_STREAM, false, F2Live<0, 8, 0, 4, -1>>(acc, T, slot, 16 * kt, wm, wn, g, tq); break;
      default: break;
    }
    F2_STAMP2(kt, 2);
  }
  cp_async_wait<0>();
  {
    double* M21 = M + o21;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int r = (2 * i + wm) * 8 + g;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = (4 * j + wn) * 8 + 2 * tq;
        *reinterpret_cast<double2*>(M21 + (long long)r * ld + c) = make_double2(-acc[i][j][0], -acc[i][j][1]);
      }
    }
  }
  F2_STAMP(9);
}

}  // namespace ffgp
