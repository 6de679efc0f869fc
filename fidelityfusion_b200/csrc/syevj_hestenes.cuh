// One-sided (Hestenes) Jacobi eigensolver for small symmetric matrices, n <= 128: the per-mode `eigh` of the Kronecker
// GP (reference MFGP_ver2023May/base_gp/hogp.py:18-22, hogp_simple.py:15-19: torch.linalg.eigh(K, UPLO='U')).
//
// Why one-sided (profiles/r01_kron_bench_v4.txt, NOTES_r01.md): the two-sided kernel (syevj_cluster_kernel) updates
// rows AND columns of A in shared memory every step - ~16 eight-byte shared accesses per 2x2 block, 4.4 us per
// round-robin step at n = 128, 3.8-4.5 ms per solve against 2.1 ms for cuSOLVER's syevd on the same box.  Here a column
// pair lives in the REGISTERS of one warp (4 doubles per lane per column at n = 128), a step is three warp-shuffle dot
// products + one plane rotation, and the only memory traffic is the tournament exchange: every warp hands one column
// to each neighbour through (distributed) shared memory and signals it with a release-stamp (no barrier per step).
//
// Algorithm.  A' = A + sigma I, sigma chosen so that A' is symmetric positive definite and well conditioned.  With
// B = min(||A||_inf, ||A||_F) >= rho(A): first attempt sigma = B / 8 - enough for a positive semi-definite A (kernel
// matrices: eigenvalues down to 1e-17 ||A||), cond(A') <= 9 and ||A'|| <= 1.125 B, which is what the backward error
// eps ||A'|| scales with; the result is VERIFIED (every singular value >= sigma / 2 and equal to its signed Rayleigh
// quotient) and only if that fails - an indefinite input - the solve is repeated with sigma = 2 B (cond <= 3 whatever
// the inertia of A).  Orthogonalise the columns of G = A' by plane rotations from the right (G <- G J): at convergence
// G = A' V = V Lambda', so lambda_i = ||g_i|| - sigma and v_i = g_i / ||g_i|| - the eigenvectors come out of G itself,
// no V is accumulated (cond <= 3 makes the normalisation as accurate as the columns), which halves the registers, the
// flops and the exchange volume.  Absolute accuracy O(eps ||A'||) <= 3 eps ||A||_inf-ish, the LAPACK class; the
// relative accuracy of tiny eigenvalues is not preserved (nor is it by syevd) and nothing downstream needs it
// (A = kron(lambda) + 1/beta, hogp.py:173-177).
//
// Layout: P = ceil(n/2) "processors" = warps, 8 per CTA, CTAs of one matrix form a cluster of C = 2, 4 or 8
// (n <= 32, 64, 128).  Tournament ordering of Brent & Luk: processor k holds (top_k, bot_k); after every step
// top_0 stays, bot_0 -> top_1, top_k -> top_{k+1}, bot_k -> bot_{k-1}, top_{P-1} -> bot_{P-1}: all n (n-1) / 2 pairs
// meet once per sweep of n - 1 steps, and every column moves to a NEIGHBOUR, so only the two boundary warps of a CTA
// write into another CTA's shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gemm_dmma.cuh"

namespace ffgp {

constexpr int HJ_WARPS = 8;
constexpr int HJ_THREADS = HJ_WARPS * 32;
constexpr int HJ_MAX_N = 128;
__device__ int g_hj_last_sweeps;      // debug: sweeps of the most recent solve (ffgp_debug_last_eigh_sweeps)

__device__ __forceinline__ void hj_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t hj_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t hj_map(const void* local, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void hj_st_f64(uint32_t addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void hj_st_s32(uint32_t addr, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t hj_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Remote store that signals the receiver: the 8 bytes land in the destination CTA's shared memory and are counted on
// the destination's mbarrier (complete_tx) - data and "it is there" travel together, no fence on either side.
__device__ __forceinline__ void hj_st_async_f64(uint32_t cluster_addr, double v, uint32_t cluster_mbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];"
               ::"r"(cluster_addr), "d"(v), "r"(cluster_mbar) : "memory");
}
__device__ __forceinline__ void hj_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void hj_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait (a protocol bug must surface as a trap, never as a hung GPU).
__device__ __forceinline__ void hj_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
// double-double accumulate:  (h, l) += a * b   (TwoProduct by FMA, TwoSum)
__device__ __forceinline__ void hj_dd_fma(double& h, double& l, double a, double b) {
  const double p = a * b, pe = fma(a, b, -p);
  const double s = h + p, bb = s - h;
  const double se = (h - (s - bb)) + (p - bb);
  h = s; l += se + pe;
}
// (h, l) += (h2, l2)
__device__ __forceinline__ void hj_dd_add(double& h, double& l, double h2, double l2) {
  const double s = h + h2, bb = s - h;
  const double se = (h - (s - bb)) + (h2 - bb);
  h = s; l += se + l2;
}
__device__ __forceinline__ double hj_max_nan(double a, double b) { return (a > b || a != a) ? a : b; }   // NaN wins
__device__ __forceinline__ double hj_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// shared memory of one CTA: inbox [HJ_WARPS][2 slots: top, bot][2 parities][R] | lam [128] | red [64] | flags [4] | mbarriers [32]
template <int NPL>
constexpr size_t hj_smem_bytes() {
  return ((size_t)HJ_WARPS * 2 * 2 * (32 * NPL) + 128 + 64) * sizeof(double) + 4 * sizeof(int) + (size_t)HJ_WARPS * 4 * sizeof(unsigned long long);
}

template <int NPL>
__global__ void __launch_bounds__(HJ_THREADS, 1)
syevj_hestenes_kernel(const double* __restrict__ Ain, int n, int C, double* __restrict__ w, double* __restrict__ V,
                      int* __restrict__ info, int max_sweeps) {
  extern __shared__ __align__(16) double hj_sm[];
  constexpr int R = 32 * NPL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = hj_cluster_rank();
  const int b = blockIdx.x / C;
  const int m = (n + 1) & ~1, P = m / 2;
  const int g = (int)rank * HJ_WARPS + warp;                   // processor index
  const bool active = g < P;
  double* inbox = hj_sm;                                       // [(warp * 2 + slot) * 2 + parity][R]
  double* lam = inbox + (size_t)HJ_WARPS * 2 * 2 * R;          // [128] eigenvalue of column position q (all CTAs hold a copy)
  double* red = lam + 128;                                     // [64] per-processor scratch
  int* flag = reinterpret_cast<int*>(red + 64);                // [3] "a rotation happened in sweep s" at s % 3 (+1 pad)
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(flag + 4);   // [HJ_WARPS][2 slots][2 parities]
  const double* A0 = Ain + (long long)b * n * n;

  if (tid < 4) flag[tid] = 0;
  if (tid < 4 * HJ_WARPS) hj_mbar_init(hj_smem_u32(&mbar[tid]), 1);
  if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  hj_cluster_sync();      // every CTA of the cluster has started and initialised its barriers before any remote access

  // ---- norms: every processor publishes max |column sum| and sum of squares of its two columns to every CTA
  double top[NPL], bot[NPL];
  const int ct = 2 * g, cb = 2 * g + 1;
  auto load_columns = [&](double sigma) {
    // processor g starts with columns 2g (top) and 2g + 1 (bot) of the symmetrised matrix (UPPER triangle is the
    // source, torch.linalg.eigh(K, UPLO='U')) + sigma I; rows >= n and the dummy column of an odd n are zero
#pragma unroll
    for (int j = 0; j < NPL; j++) {
      const int r = lane + 32 * j;
      double vt = 0.0, vb = 0.0;
      if (active && r < n) {
        if (ct < n) vt = A0[(long long)min(r, ct) * n + max(r, ct)] + (r == ct ? sigma : 0.0);
        if (cb < n) vb = A0[(long long)min(r, cb) * n + max(r, cb)] + (r == cb ? sigma : 0.0);
      }
      top[j] = vt; bot[j] = vb;
    }
  };
  load_columns(0.0);
  double st = 0.0, sb = 0.0, sq = 0.0;
#pragma unroll
  for (int j = 0; j < NPL; j++) {
    st += fabs(top[j]); sb += fabs(bot[j]);
    sq = fma(top[j], top[j], sq); sq = fma(bot[j], bot[j], sq);
  }
  st = hj_warp_sum(st); sb = hj_warp_sum(sb); sq = hj_warp_sum(sq);
  if (active && lane < C) {
    hj_st_f64(hj_map(&red[g], (uint32_t)lane), hj_max_nan(st, sb));
    hj_st_f64(hj_map(&lam[g], (uint32_t)lane), sq);
  }
  hj_cluster_sync();
  double norm_inf = 0.0, norm_fro = 0.0;
  for (int k = lane; k < P; k += 32) { norm_inf = hj_max_nan(norm_inf, red[k]); norm_fro += lam[k]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    norm_inf = hj_max_nan(norm_inf, __shfl_xor_sync(0xffffffffu, norm_inf, o));
    norm_fro += __shfl_xor_sync(0xffffffffu, norm_fro, o);
  }
  norm_fro = sqrt(norm_fro);
  const double bound = (norm_fro < norm_inf) ? norm_fro : norm_inf;       // >= spectral radius (NaN-propagating)
  const bool bad_input = !(bound < 1.0e300);                   // inf / nan input: report, do not iterate
  const bool zero_input = bound == 0.0;
  hj_cluster_sync();                                           // lam / red are reused below

  // ---- exchange wiring (fixed for the whole solve).  slot 0 = "new top" inbox, slot 1 = "new bot" inbox of a warp.
  int dt_g = -1, dt_slot = 0, db_g = -1, db_slot = 0;          // destination processor / slot of my top and my bot
  if (active && P > 1) {
    if (g == P - 1) { dt_g = g; dt_slot = 1; }                 // top_{P-1} -> bot_{P-1}
    else if (g >= 1) { dt_g = g + 1; dt_slot = 0; }            // top_k -> top_{k+1}
    if (g == 0) { db_g = 1; db_slot = 0; }                     // bot_0 -> top_1
    else { db_g = g - 1; db_slot = 1; }                        // bot_k -> bot_{k-1}
  }
  auto slot_addr = [&](int pg, int slot) -> uint32_t {
    const int pw = pg % HJ_WARPS;
    return hj_map(inbox + (size_t)((pw * 2 + slot) * 2) * R, (uint32_t)(pg / HJ_WARPS));
  };
  const uint32_t dt_addr = dt_g >= 0 ? slot_addr(dt_g, dt_slot) : 0u;
  const uint32_t db_addr = db_g >= 0 ? slot_addr(db_g, db_slot) : 0u;
  const double* in_top = inbox + (size_t)((warp * 2 + 0) * 2) * R;
  const double* in_bot = inbox + (size_t)((warp * 2 + 1) * 2) * R;
  const bool recv_top = active && P > 1 && g >= 1;             // processor 0 keeps its top

  const double HUGE_L = 1.0e308;
  bool converged = zero_input || bad_input;
  int sweep = 0;
  unsigned int nstep = 0;
  double lt = 0.0, lb = 0.0;
  bool dummy_t = false, dummy_b = false;
  for (int attempt = 0; attempt < 2; attempt++) {
  const double sigma = attempt == 0 ? 0.125 * bound : 2.0 * bound;
  load_columns(sigma);
  converged = zero_input || bad_input;
  const double tol2 = (double)m * 4.930380657631324e-32;       // (sqrt(m) eps)^2, the dgesvj threshold
  // Neighbour signalling instead of a cluster barrier per step: a column only ever moves to a NEIGHBOURING processor,
  // so a warp waits for exactly the two warps it receives from.  The sender writes the column into the receiver's inbox
  // with st.async, whose bytes are counted on the receiver's mbarrier of that (slot, parity); the receiver arms the
  // barrier with expect_tx = one column and waits on its phase.  (First versions: barrier.cluster per step 1.2 us / step;
  // release-stamp + acquire-polling at cluster scope 7.8 us / step - every cluster-scope fence costs ~1.5 us.)
  // Flow control is implicit: the exchange between neighbours is symmetric (k sends to k+1 and k+1 sends to k every
  // step), so a sender can be at most one step ahead of its receiver and parity p is rewritten only after the receiver
  // has sent - i.e. has consumed - the step before.  One cluster barrier per SWEEP remains (convergence vote).
  const uint32_t my_bar = hj_smem_u32(&mbar[warp * 4]);        // + 16 * slot + 8 * parity
  auto bar_addr = [&](int pg, int slot) -> uint32_t {
    return hj_map(&mbar[(pg % HJ_WARPS) * 4 + slot * 2], (uint32_t)(pg / HJ_WARPS));
  };
  const uint32_t dt_bar = dt_g >= 0 ? bar_addr(dt_g, dt_slot) : 0u;
  const uint32_t db_bar = db_g >= 0 ? bar_addr(db_g, db_slot) : 0u;
  for (sweep = 0; sweep < max_sweeps && !converged; sweep++) {
    bool rotated = false;
    if (active) {
      for (int step = 0; step < m - 1; step++, nstep++) {
        const unsigned int par = nstep & 1u;
        double a = 0.0, bb = 0.0, c = 0.0;
#pragma unroll
        for (int j = 0; j < NPL; j++) {
          a = fma(top[j], top[j], a);
          bb = fma(bot[j], bot[j], bb);
          c = fma(top[j], bot[j], c);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {                     // three interleaved butterflies
          a += __shfl_xor_sync(0xffffffffu, a, o);
          bb += __shfl_xor_sync(0xffffffffu, bb, o);
          c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (c * c > tol2 * a * bb) {                           // (false for a zero column: the dummy never rotates)
          rotated = true;
          // tan of the rotation that zeroes g_top . g_bot:  t = sign(d) 2c / (|d| + sqrt(d^2 + 4 c^2)),  d = bb - a.
          // rsqrt / reciprocal instead of IEEE sqrt and divide: the angle only has to be good to a few ulp (Jacobi
          // corrects itself), and [[cs, -sn], [sn, cs]] keeps the two columns orthogonal whatever cs^2 + sn^2 is
          const double d = bb - a, c2 = 2.0 * c;
          const double h2 = fma(d, d, c2 * c2);
          const double den = fabs(d) + h2 * rsqrt(h2);
          const double t = (d >= 0.0 ? c2 : -c2) * __drcp_rn(den);
          const double cs = rsqrt(fma(t, t, 1.0)), sn = cs * t;
#pragma unroll
          for (int j = 0; j < NPL; j++) {
            const double x = top[j], y = bot[j];
            top[j] = fma(cs, x, -sn * y);
            bot[j] = fma(sn, x, cs * y);
          }
        }
        if (P > 1) {
          const uint32_t ph = (nstep >> 1) & 1u;               // phase of the (slot, parity) barriers in this step
          if (lane == 0) {                                     // arm my two inboxes for this step
            if (recv_top) hj_mbar_expect_tx(my_bar + 8u * par, (uint32_t)(R * 8));
            hj_mbar_expect_tx(my_bar + 16u + 8u * par, (uint32_t)(R * 8));
          }
          if (dt_g >= 0) {
#pragma unroll
            for (int j = 0; j < NPL; j++)
              hj_st_async_f64(dt_addr + (uint32_t)((par * R + lane + 32 * j) * 8), top[j], dt_bar + 8u * par);
          }
#pragma unroll
          for (int j = 0; j < NPL; j++)
            hj_st_async_f64(db_addr + (uint32_t)((par * R + lane + 32 * j) * 8), bot[j], db_bar + 8u * par);
          if (recv_top) {
            hj_mbar_wait(my_bar + 8u * par, ph);
#pragma unroll
            for (int j = 0; j < NPL; j++) top[j] = in_top[par * R + lane + 32 * j];
          }
          hj_mbar_wait(my_bar + 16u + 8u * par, ph);
#pragma unroll
          for (int j = 0; j < NPL; j++) bot[j] = in_bot[par * R + lane + 32 * j];
        }
      }
      if (rotated && lane < C) hj_st_s32(hj_map(&flag[sweep % 3], (uint32_t)lane), 1);
    }
    hj_cluster_sync();
    converged = flag[sweep % 3] == 0;
    if (tid == 0) flag[(sweep + 2) % 3] = 0;                   // next written in sweep + 2: at least one barrier away
  }

  // ---- eigenvalues.  v_i = g_i / ||g_i||; lambda_i = Rayleigh quotient v_i^T A v_i with the ORIGINAL matrix, products
  // and sums carried in double-double (TwoProduct / TwoSum): the quotient's error is second order in the eigenvector
  // error (~1e-28 ||A||) and its evaluation error is eps |lambda|, so the eigenvalues are accurate to the last digits of
  // the input matrix instead of the eps ||A'|| of (||g_i|| - sigma).  This matters downstream: A = kron(lambda) + 1/beta
  // multiplies an absolute error in one mode's lambda by the other modes' largest eigenvalues (hogp.py:173-177).
  double nt = 0.0, nb = 0.0;
#pragma unroll
  for (int j = 0; j < NPL; j++) { nt = fma(top[j], top[j], nt); nb = fma(bot[j], bot[j], nb); }
  nt = sqrt(hj_warp_sum(nt)); nb = sqrt(hj_warp_sum(nb));
  dummy_t = zero_input ? (ct >= n) : !(nt > 0.0); dummy_b = zero_input ? (cb >= n) : !(nb > 0.0);
  {
    const double it = nt > 0.0 ? 1.0 / nt : 0.0, ib = nb > 0.0 ? 1.0 / nb : 0.0;
#pragma unroll
    for (int j = 0; j < NPL; j++) { top[j] *= it; bot[j] *= ib; }
  }
  lt = 0.0; lb = 0.0;
  if (active && !zero_input && !bad_input) {
    double* vsh = inbox + (size_t)(warp * 4) * R;              // this warp's inbox (idle now): v_top [R], v_bot [R]
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NPL; j++) { vsh[lane + 32 * j] = top[j]; vsh[R + lane + 32 * j] = bot[j]; }
    __syncwarp();
    double yt[NPL], yte[NPL], yb[NPL], ybe[NPL];               // (A v)_r for the rows r = lane + 32 j, hi + lo parts
#pragma unroll
    for (int j = 0; j < NPL; j++) { yt[j] = yte[j] = yb[j] = ybe[j] = 0.0; }
    for (int c = 0; c < n; c++) {
      const double vt = vsh[c], vb = vsh[R + c];
#pragma unroll
      for (int j = 0; j < NPL; j++) {
        const int r = lane + 32 * j;
        const double arc = r < n ? A0[(long long)min(r, c) * n + max(r, c)] : 0.0;
        hj_dd_fma(yt[j], yte[j], arc, vt);
        hj_dd_fma(yb[j], ybe[j], arc, vb);
      }
    }
    double st_h = 0.0, st_l = 0.0, sb_h = 0.0, sb_l = 0.0;
#pragma unroll
    for (int j = 0; j < NPL; j++) {
      hj_dd_fma(st_h, st_l, yt[j], top[j]); st_l = fma(yte[j], top[j], st_l);
      hj_dd_fma(sb_h, sb_l, yb[j], bot[j]); sb_l = fma(ybe[j], bot[j], sb_l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      hj_dd_add(st_h, st_l, __shfl_xor_sync(0xffffffffu, st_h, o), __shfl_xor_sync(0xffffffffu, st_l, o));
      hj_dd_add(sb_h, sb_l, __shfl_xor_sync(0xffffffffu, sb_h, o), __shfl_xor_sync(0xffffffffu, sb_l, o));
    }
    // (the double-double butterfly is not bitwise symmetric between partner lanes: take lane 0's value everywhere, the
    // ranks below compare it with the copy lane 0 publishes)
    lt = __shfl_sync(0xffffffffu, st_h + st_l, 0); lb = __shfl_sync(0xffffffffu, sb_h + sb_l, 0);
    // verification of the shift: A' is positive definite and well conditioned iff every singular value ||g_i|| is
    // >= sigma / 2 and equals the signed Rayleigh quotient of A' (a negative or +/- paired eigenvalue of A' breaks it)
    if (attempt == 0 && converged) {
      const double tolv = 1.0e-8 * bound;
      const bool ok_t = dummy_t || (nt >= 0.5 * sigma && fabs(lt + sigma - nt) <= tolv);
      const bool ok_b = dummy_b || (nb >= 0.5 * sigma && fabs(lb + sigma - nb) <= tolv);
      if (!(ok_t && ok_b) && lane < C) hj_st_s32(hj_map(&flag[3], (uint32_t)lane), 1);
    }
  }
  if (dummy_t) lt = HUGE_L;
  if (dummy_b) lb = HUGE_L;
  if (active && lane < C) {
    hj_st_f64(hj_map(&lam[2 * g], (uint32_t)lane), lt);
    hj_st_f64(hj_map(&lam[2 * g + 1], (uint32_t)lane), lb);
  }
  hj_cluster_sync();
  const bool retry = attempt == 0 && flag[3] != 0;
  if (retry) {                                                 // indefinite input: once more with the safe shift
    hj_cluster_sync();                                         // (everyone has read flag[3] and lam)
    continue;
  }
  if (active) {
    int rt = 0, rb = 0;
    for (int q = lane; q < m; q += 32) {
      const double l = lam[q];
      rt += (l < lt) || (l == lt && q < 2 * g);
      rb += (l < lb) || (l == lb && q < 2 * g + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { rt += __shfl_xor_sync(0xffffffffu, rt, o); rb += __shfl_xor_sync(0xffffffffu, rb, o); }
    double* Vb = V + (long long)b * n * n;
    if (rt < n) {
      if (lane == 0) w[(long long)b * n + rt] = lt;
#pragma unroll
      for (int j = 0; j < NPL; j++) {
        const int r = lane + 32 * j;
        if (r < n) Vb[(long long)r * n + rt] = zero_input ? (r == ct ? 1.0 : 0.0) : top[j];
      }
    }
    if (rb < n) {
      if (lane == 0) w[(long long)b * n + rb] = lb;
#pragma unroll
      for (int j = 0; j < NPL; j++) {
        const int r = lane + 32 * j;
        if (r < n) Vb[(long long)r * n + rb] = zero_input ? (r == cb ? 1.0 : 0.0) : bot[j];
      }
    }
  }
  if (!retry) break;
  }   // attempt
  if (rank == 0 && tid == 0 && (!converged || bad_input)) info[b] = 1;
  if (rank == 0 && tid == 0 && b == 0) g_hj_last_sweeps = sweep;
}

// cluster size for n: P = ceil(n/2) processors, 8 per CTA, rounded up to a power of two
inline int hj_cluster_size(int n) {
  const int P = (n + 1) / 2;
  const int need = (P + HJ_WARPS - 1) / HJ_WARPS;
  int C = 2;            // at least 2: st.async (shared::cluster + remote mbarrier completion) needs a real cluster; for
  while (C < need) C *= 2;   // n <= 16 the second CTA holds no processor and only takes part in the cluster barriers
  return C;
}

template <int NPL>
inline cudaError_t hj_launch_npl(const double* A, int n, int batch, double* w, double* V, int* info, int max_sweeps,
                                 cudaStream_t st) {
  const int C = hj_cluster_size(n);
  static PerDeviceOnce once;
  bool& attr_set = *once.slot();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(syevj_hestenes_kernel<NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)hj_smem_bytes<NPL>());
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(batch * C), 1, 1);
  cfg.blockDim = dim3(HJ_THREADS, 1, 1);
  cfg.dynamicSmemBytes = hj_smem_bytes<NPL>();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, syevj_hestenes_kernel<NPL>, A, n, C, w, V, info, max_sweeps);
}

inline cudaError_t hj_launch(const double* A, int n, int batch, double* w, double* V, int* info, int max_sweeps,
                             cudaStream_t st) {
  const int npl = (n + 31) / 32;
  switch (npl) {
    case 1: return hj_launch_npl<1>(A, n, batch, w, V, info, max_sweeps, st);
    case 2: return hj_launch_npl<2>(A, n, batch, w, V, info, max_sweeps, st);
    case 3: return hj_launch_npl<3>(A, n, batch, w, V, info, max_sweeps, st);
    case 4: return hj_launch_npl<4>(A, n, batch, w, V, info, max_sweeps, st);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace ffgp
