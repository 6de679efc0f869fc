// Warp-specialised FP64 tensor-pipe GEMM for sm_100a: TMA-staged operand panels, mbarrier ring, DMMA consumers.
//
// Why (profiles/r01_ncu_gemm_v2_stalls.txt): the cp.async kernel in gemm_dmma.cuh kept the DMMA pipe 87 % busy - every
// k-block all eight MMA warps met at one __syncthreads and then spent ~200 issue slots each on cp.async address
// arithmetic while the pipe drained.  Here the MMA warps never compute a global address and never meet at a CTA
// barrier:
//   * warp 8 (one elected lane of the producer warpgroup) is the producer: per k-block it waits for the stage's `empty` mbarrier, posts the
//     byte count on the `full` mbarrier and issues two cp.async.bulk.tensor (TMA) loads - one 128 x 16 panel of A,
//     one of B - which land in shared memory in a bank-conflict-free layout chosen by the tensor map:
//       k-major operand  (p contiguous):  4-D map (k, row, inner batch, outer batch), box 16 x 128, SWIZZLE_128B
//       m-major operand  (row contiguous): 5-D "panel view" (row%8, k, row/8, inner, outer) with strides
//                                          (8, ld*8, 64, ..) bytes, box 8 x 16 x 16 -> smem [16 panels][16 k][8 rows]
//     so every m8n8k4 fragment load (8 rows x 4 k) touches 256 contiguous-in-bank bytes = 2 wavefronts, the minimum;
//   * warps 0-7 (2 x 4, warp tile 64 x 32 = 32 DMMA per k4 step) wait on `full`, keep fragments double-buffered in
//     registers one k4 step ahead (the wait for the next stage hides behind 32 DMMAs) and release the stage with one
//     mbarrier arrive per warp.
// 6 stages x 32 KB.  One tile per CTA, heavy tiles first: the hardware CTA scheduler is the load balancer and leaves
// SMs to the high-priority look-ahead stream as tiles retire (a persistent grid would starve it).
// FP64 has no tcgen05/TMEM path on sm_100a: DMMA.8x8x4 (mma.sync.m8n8k4.f64) with register operands is the FP64
// tensor pipe; TMA + mbarrier are the Blackwell/Hopper pieces that apply.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include "gemm_dmma.cuh"

namespace ffgp {

// Stage release of the operand ring.  History: an mbarrier arrive does not wait for the warp's outstanding ld.shared, and
// ptxas is free to sink fragment loads and hoist the arrive, so "arrive after the loads were ISSUED" is a race with the
// TMA refill of the stage (generic-proxy reads vs async-proxy writes).  It showed once the ring stayed full and the
// load/store queue was backed up (persistent grid + 64 global loads of C per thread at the start of a tile): one warp's
// last-loaded A / B fragments were overwritten before they were served, ~1 tile in 500 wrong (tools/diag_stage_release.py).
//   TG_RELEASE_LATE = 0  (round 1) early release, one k4 step before the end of the stage, kept correct by making the
//                        barrier ADDRESS data-dependent on every fragment read from the stage (`rel & p.zero`, a run-time
//                        zero): works, but it is a scheduling trick, not a memory-model guarantee;
//   TG_RELEASE_LATE = 1  (default) release after the DMMAs of the stage's last k4 step have issued: a DMMA cannot issue
//                        before the ld.shared feeding its operands have returned, a warp issues in order and the
//                        compiler keeps the order of the asm volatile statements - provably ordered, 24 LOP3 per k4 step
//                        fewer, and the stage is handed back ~500 clk later out of a 6-stage ring
//                        (A/B on one box: profiles/r02_c5_experiments.txt).
#ifndef TG_RELEASE_LATE
#define TG_RELEASE_LATE 1
#endif
constexpr int TG_BM = 128, TG_BN = 128, TG_BK = 16, TG_STAGES = 6;
constexpr int TG_CONSUMER_WARPS = 8;
constexpr int TG_THREADS = (TG_CONSUMER_WARPS + 4) * 32;             // 2 consumer warpgroups + 1 producer warpgroup
// Registers are a per-SM-sub-partition resource (16384 each): 12 warps launch with <= 168 registers per thread, then
// setmaxnreg moves the producer warpgroup's share to the MMA warps (40 / 232), which need 128 accumulator registers
// plus two fragment buffers.
constexpr int TG_REGS_PRODUCER = 40, TG_REGS_CONSUMER = 232;
constexpr int TG_PANEL_BYTES = TG_BM * TG_BK * 8;                       // 16 KB per operand per stage
constexpr int TG_STAGE_BYTES = 2 * TG_PANEL_BYTES;
constexpr int TG_TILE_RING = 8;                                         // > TG_STAGES: the producer is at most TG_STAGES tiles ahead
constexpr size_t TG_SMEM_BYTES = (size_t)TG_STAGES * TG_STAGE_BYTES + 1024 /*align slack*/ + 128 /*barriers*/ + 4 * TG_TILE_RING + 16 /*stagger flag*/;

// Tile shape by the number of warp COLUMNS WN (warp tile 64 x 32 either way, same per-warp code and register budget):
//   WN = 4: 128 x 128 tile, 8 MMA warps + producer warpgroup = 384 threads, 6 x 32 KB stages, ONE CTA per SM
//           (168 registers at launch, setmaxnreg 40 / 232).  Long-K tiles: 0.99 of the DGEMM rate at K = 8192.
//   WN = 2: 128 x 64 tile, 4 MMA warps + producer warpgroup = 256 threads, 4 x 24 KB stages, TWO CTAs per SM
//           (128 registers at launch, setmaxnreg 24 / 232): the tile hand-over of one CTA (epilogue stores, decode,
//           accumulator init, first barrier) hides behind the other CTA's DMMAs.  For the 128/256-deep tiles of the
//           batched factorisation, where that hand-over is 15-25 % of a tile (profiles/r02_gemm_small_k_tiles.txt),
//           and for 64-wide right-hand sides (the posterior's V = M K*).  Not for lower_only launches (square tiles).
template <int WN>
struct TgCfg {
  static constexpr int BN = 32 * WN;
  static constexpr int CW = 2 * WN;                                     // MMA warps
  static constexpr int THREADS = (CW + 4) * 32;
  static constexpr int STAGES = WN == 4 ? TG_STAGES : 4;
  static constexpr int B_PANEL_BYTES = BN * TG_BK * 8;
  static constexpr int STAGE_BYTES = TG_PANEL_BYTES + B_PANEL_BYTES;
  static constexpr int REGS_PRODUCER = WN == 4 ? TG_REGS_PRODUCER : 24;
  static constexpr int CTAS_PER_SM = WN == 4 ? 1 : 2;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 128 + 4 * TG_TILE_RING + 16;
};

struct TmaGemmParams {
  double* C;
  int ldc;
  long long sC, iC;
  int M, N, K, inner;
  double alpha, beta;
  int lower_only, kmode, heavy_first;
  int zero;                  // always 0 (a run-time value the compiler cannot fold, see the stage release)
  int tiles, total;          // tiles per problem, work items of the launch (tiles x problems); grid.x <= total
  int stagger;               // persistent grid only: k-steps by which the second warp row starts behind the first (0: off)
  unsigned int* sched;       // NULL: one work item per CTA (grid.x == total).  Else {next, done} counters, both zero at
                             // launch and reset by the last CTA: CTAs fetch further work items as they finish
};

__device__ __forceinline__ uint32_t tg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tg_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tg_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tg_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must surface as a trap, never as a hung GPU.
__device__ __forceinline__ void tg_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  if (done) return;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tg_tma_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tg_tma_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6,%7}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ double tg_lds(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// Compile-time set of live 8x8 accumulator fragments of a consumer warp in one k-step (see the kernel):
// i in [ILO, IHI), j in [JLO, JHI), and for ROLE = 4 wm + wn >= 0 only fragments on/below the tile diagonal.
template <int ILO, int IHI, int JLO, int JHI, int ROLE>
struct TgLive {
  __device__ static constexpr bool live(int i, int j) {
    return i >= ILO && i < IHI && j >= JLO && j < JHI && (ROLE < 0 || 2 * i + (ROLE >> 2) - 4 * j - (ROLE & 3) >= 0);
  }
};

// Work item w of a launch -> problem (zo, zi), tile (ti, tj) and the K range of that tile (same ordering rules as
// gemm_dmma_kernel).  Items are numbered tile-fastest, so the tiles of one problem are in flight together and share
// their operand panels through L2.
struct TgTile { int ti, tj, k_lo, k_hi, zo, zi; };
template <int BN>
__device__ __forceinline__ TgTile tg_decode(const TmaGemmParams& p, int w) {
  TgTile r;
  const int z = w / p.tiles;
  int t = w - z * p.tiles;
  const int tiles_m = p.M / TG_BM, tiles_n = p.N / BN;
  if (p.lower_only) {
    if (p.heavy_first && p.kmode != K_GE_ROW) t = p.tiles - 1 - t;
    r.ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (r.ti * (r.ti + 1) / 2 > t) --r.ti;
    while ((r.ti + 1) * (r.ti + 2) / 2 <= t) ++r.ti;
    r.tj = t - r.ti * (r.ti + 1) / 2;
  } else {
    if (p.heavy_first && (p.kmode == K_LE_ROW || p.kmode == K_LE_COL)) t = p.tiles - 1 - t;
    if (p.kmode == K_LE_COL || p.kmode == K_GE_COL) { r.tj = t / tiles_m; r.ti = t - r.tj * tiles_m; }
    else { r.ti = t / tiles_n; r.tj = t - r.ti * tiles_n; }
  }
  const int i0 = r.ti * TG_BM, j0 = r.tj * BN;
  r.k_lo = 0; r.k_hi = p.K;
  if (p.kmode == K_LE_ROW) r.k_hi = min(p.K, i0 + TG_BM);
  else if (p.kmode == K_LE_COL) r.k_hi = min(p.K, j0 + BN);
  else if (p.kmode == K_GE_COL) r.k_lo = j0;
  else if (p.kmode == K_GE_ROW) r.k_lo = i0;
  r.zo = z / p.inner; r.zi = z - r.zo * p.inner;
  return r;
}

#ifdef FFGP_TG_TRACE
// Per-tile timeline of consumer warp 0 of CTA 0 (tools/tg_trace.py): clock64 at tile start | accumulators initialised |
// first stage arrived + first fragments requested | main loop done | epilogue done | next tile id known.
__device__ long long g_tg_trace[2 * 32 * 6];           // [warp 0 | warp 4 (same SM sub-partition)][tile][stamp]
#define TG_STAMP(n, s) do { if (blockIdx.x == 0 && (warp & 3) == 0 && lane == 0 && (n) < 32) g_tg_trace[((warp >> 2) * 32 + (n)) * 6 + (s)] = clock64(); } while (0)
__device__ long long g_tg_cta[4096 * 4];                // non-persistent launches: per CTA {tile start, main loop start, main loop end, end}
#define TG_CTA(s) do { if (!p.sched && blockIdx.x < 4096 && warp == 0 && lane == 0) g_tg_cta[blockIdx.x * 4 + (s)] = clock64(); } while (0)
#else
#define TG_STAMP(n, s) do { } while (0)
#define TG_CTA(s) do { } while (0)
#endif

// A_KMAJ: A(i,p) = A[i*lda + p]  else  A(i,p) = A[p*lda + i];   B_KMAJ: B(p,j) = B[j*ldb + p]  else  B[p*ldb + j]
template <bool A_KMAJ, bool B_KMAJ, int WN = 4>
__global__ void __launch_bounds__(TgCfg<WN>::THREADS, TgCfg<WN>::CTAS_PER_SM)
gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TmaGemmParams p) {
  using Cfg = TgCfg<WN>;
  constexpr int TG_BN = Cfg::BN, TG_STAGES = Cfg::STAGES, TG_STAGE_BYTES = Cfg::STAGE_BYTES, TG_CONSUMER_WARPS = Cfg::CW;
  extern __shared__ unsigned char tg_smem_raw[];
  const uint32_t smem_base = (tg_smem_u32(tg_smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B panels need 1024-B alignment
  const uint32_t bar_base = smem_base + TG_STAGES * TG_STAGE_BYTES;           // full[s] at +8s, empty[s] at +64+8s
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  volatile int* tile_ring = reinterpret_cast<volatile int*>(tg_smem_raw + (smem_base - tg_smem_u32(tg_smem_raw)) +
                                                            TG_STAGES * TG_STAGE_BYTES + 128);

  volatile int* stagger_go = tile_ring + TG_TILE_RING;
  if (tid == 0) {
    *stagger_go = 0;
    for (int s = 0; s < TG_STAGES; s++) {
      tg_mbar_init(bar_base + 8 * s, 1);                           // full: one arrive.expect_tx by the producer
      tg_mbar_init(bar_base + 64 + 8 * s, TG_CONSUMER_WARPS);      // empty: one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= TG_CONSUMER_WARPS) {
    // ===================== TMA producer warpgroup (one elected lane works) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_PRODUCER));
    if (warp == TG_CONSUMER_WARPS && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
      // Work items: the CTA's own (blockIdx.x), then - persistent grid - whatever the launch-wide counter hands out
      // next, so SMs take tiles as they free up exactly like the hardware CTA scheduler would, minus the per-CTA launch,
      // barrier-init and first-load latency.  The id of every further tile is published to the MMA warps through
      // tile_ring (written before the expect_tx arrive of the tile's first stage, read after the wait on it); -1 ends.
      int it = 0;                                   // k-steps issued so far: the stage ring runs on across tiles
      int w = blockIdx.x;
      for (int n = 0;; n++) {
        unsigned int fetched = 0;
        if (p.sched) fetched = atomicAdd(p.sched, 1u);      // next item, requested before this one's loads are issued
        const TgTile tl = tg_decode<TG_BN>(p, w);
        const int i0 = tl.ti * TG_BM, j0 = tl.tj * TG_BN;
        const int KT = (tl.k_hi - tl.k_lo) / TG_BK;
        for (int kt = 0; kt < KT; kt++, it++) {
          const int s = it % TG_STAGES;
          const uint32_t full = bar_base + 8 * s, empty = bar_base + 64 + 8 * s;
          if (it >= TG_STAGES) tg_mbar_wait(empty, ((it / TG_STAGES) - 1) & 1);
          if (kt == 0 && n > 0) tile_ring[n % TG_TILE_RING] = w;
          tg_mbar_expect_tx(full, TG_STAGE_BYTES);
          const uint32_t dstA = smem_base + s * TG_STAGE_BYTES, dstB = dstA + TG_PANEL_BYTES;
          const int k0 = tl.k_lo + kt * TG_BK;
          if (A_KMAJ) tg_tma_4d(dstA, &mapA, full, k0, i0, tl.zi, tl.zo);
          else tg_tma_5d(dstA, &mapA, full, 0, k0, i0 >> 3, tl.zi, tl.zo);
          if (B_KMAJ) tg_tma_4d(dstB, &mapB, full, k0, j0, tl.zi, tl.zo);
          else tg_tma_5d(dstB, &mapB, full, 0, k0, j0 >> 3, tl.zi, tl.zo);
        }
        if (!p.sched) break;
        const long long wn = (long long)gridDim.x + fetched;
        if (wn >= p.total) {
          // end marker: takes one stage slot (no data), the MMA warps leave without releasing it
          const int s = it % TG_STAGES;
          if (it >= TG_STAGES) tg_mbar_wait(bar_base + 64 + 8 * s, ((it / TG_STAGES) - 1) & 1);
          tile_ring[(n + 1) % TG_TILE_RING] = -1;
          tg_mbar_arrive(bar_base + 8 * s);
          // every CTA fetches exactly one item past the end: the last one to get there rearms the counters
          if (atomicAdd(p.sched + 1, 1u) == gridDim.x - 1) { p.sched[0] = 0u; p.sched[1] = 0u; __threadfence(); }
          break;
        }
        w = (int)wn;
      }
    }
    return;
  }

  // ===================== DMMA consumers =====================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TG_REGS_CONSUMER));
  // 2 x 4 warps, warp tile 64 x 32 made of CYCLICALLY assigned 8-wide fragments: m-fragment i of warp row wm covers tile
  // rows (2 i + wm) * 8 .., n-fragment j of warp column wn covers tile columns (4 j + wn) * 8 ..  Every warp then owns
  // the same share of any triangular region of the tile, so structural zeros (below) can be skipped without one warp
  // becoming the straggler.  Warps w and w + 4 share an SM sub-partition: wn is mirrored in the second warp row so the
  // sub-partitions' DMMA counts on a symmetric diagonal tile are 36/32/36/32 of 64.
  const int wm = warp / WN, wn = (warp % WN) ^ (wm ? WN - 1 : 0);
  const int g = lane >> 2, tq = lane & 3;
  constexpr int MT = 8, NT = 4;
  // per-thread fragment offsets inside a stage (bytes)
  const uint32_t swz_h = (uint32_t)((tq >> 1) ^ g);        // k-major: 16-B chunk index = (2 kk + (tq>>1)) ^ (row & 7)
  const uint32_t a_off = A_KMAJ ? (uint32_t)((wm * 8 + g) * 128 + ((tq & 1) << 3))
                                : (uint32_t)(wm * 1024 + tq * 64 + g * 8);
  const uint32_t b_off = TG_PANEL_BYTES + (B_KMAJ ? (uint32_t)((wn * 8 + g) * 128 + ((tq & 1) << 3))
                                                  : (uint32_t)(wn * 1024 + tq * 64 + g * 8));
  double af[2][MT], bf[2][NT];
  auto load_frags = [&](int buf, uint32_t stage_base, int kk) {
    const uint32_t kx = A_KMAJ || B_KMAJ ? (((uint32_t)(2 * kk) ^ swz_h) << 4) : 0u;
    const uint32_t pa = stage_base + a_off + (A_KMAJ ? kx : (uint32_t)(kk * 256));
    const uint32_t pb = stage_base + b_off + (B_KMAJ ? kx : (uint32_t)(kk * 256));
#pragma unroll
    for (int i = 0; i < MT; i++) af[buf][i] = tg_lds(pa + i * 2048);
#pragma unroll
    for (int j = 0; j < NT; j++) bf[buf][j] = tg_lds(pb + j * (WN * 1024));
  };

  const double alpha = p.alpha, beta = p.beta;
  // beta C enters through the accumulators: the tile of C is read while the first panels are in flight
  const bool init_from_c = (beta != 0.0) && (alpha != 0.0);
  const int role = wm * 4 + wn;
  double acc[MT][NT][2];
  // ring position of the NEXT k-step to consume (the stage ring runs on across tiles): stage index and phase parity are
  // carried incrementally - `it % TG_STAGES` and `it / TG_STAGES` per k-step were two multiply-high sequences in front of the
  // first DMMA of every k-step
  int cs = 0;
  uint32_t cph = 0;
  int w = blockIdx.x;
  // Stagger experiment (FFGP_STAGGER = k-steps, default 0 = off; persistent grid only).  Warps w and w + 4 share an SM
  // sub-partition; started together, both reach every tile's epilogue + accumulator init (~4.4k clk per tile) at the same
  // moment and the DMMA pipe idles.  Starting the second warp row `stagger` k-steps late keeps the rows out of phase for
  // all of the CTA's tiles (measured: the 17k-clk offset persists) - but a tile takes 38.6k clk either way
  // (profiles/r02_c5_experiments.txt): alone on its sub-partition a warp of this main loop issues a DMMA per 22.5 clk
  // (17.4 when two share the pipe), and the other row's 32 STG.128 per thread sit in the same load/store path as this
  // row's fragment loads, so the hand-over is not hidden.  Kept as a knob; the offset must stay below TG_STAGES.
  const bool staggered = p.sched != nullptr && p.stagger > 0;
  if (staggered && wm == 1) {
#pragma unroll 1
    for (uint32_t spin = 0; *stagger_go == 0; ++spin) {
      __nanosleep(64);
      if (spin > (1u << 24)) __trap();                 // a protocol bug must surface as a trap, never as a hung GPU
    }
  }
  for (int n = 0;; n++) {
  TG_STAMP(n, 0);
  TG_CTA(0);
  const TgTile tl = tg_decode<TG_BN>(p, w);
  const int ti = tl.ti, tj = tl.tj, i0 = ti * TG_BM, j0 = tj * TG_BN, k_hi = tl.k_hi;
  const int KT = (tl.k_hi - tl.k_lo) / TG_BK;
  // ---- structural zeros.  A kmode says an operand is TRIANGULAR (ffgp.h): besides shortening the K range of the tile,
  // inside the one 128-deep K block that straddles the diagonal a k4 step only touches fragments whose rows (columns)
  // reach it; and a lower_only diagonal tile is symmetric, only fragments on/below its diagonal are produced (the rest
  // of the tile is left untouched).  At N = 512 (BASELINE config 5) this removes 1/3 of the DMMAs of a factorisation.
  const bool sym_diag = p.lower_only && ti == tj;
  int tri_kt0 = -1;                                        // first k-step (of 8) of the diagonal K block, -1: none
  int tri_mode = 0;                                        // 1: A rows >= p   2: B cols >= p   3: B cols <= p   4: A rows <= p
  if (p.kmode == K_LE_ROW && k_hi == i0 + TG_BM) { tri_mode = 1; tri_kt0 = KT - TG_BM / TG_BK; }
  else if (p.kmode == K_LE_COL && k_hi == j0 + TG_BN) { tri_mode = 2; tri_kt0 = KT - TG_BN / TG_BK; }
  else if (p.kmode == K_GE_COL) { tri_mode = 3; tri_kt0 = 0; }
  else if (p.kmode == K_GE_ROW) { tri_mode = 4; tri_kt0 = 0; }
  const int sym_thr = sym_diag ? 0 : -64;                 // fragment (i, j) is produced iff 2 i + wm - 4 j - wn >= sym_thr
  double* __restrict__ Cg = p.C + (long long)tl.zo * p.sC + (long long)tl.zi * p.iC;

  if (init_from_c) {
    // C is read while the first panels are in flight: acc starts at (beta/alpha) C, the epilogue multiplies by alpha
    const double r = beta / alpha;
#pragma unroll
    for (int i = 0; i < MT; i++) {
      const int row = i0 + (2 * i + wm) * 8 + g;
#pragma unroll
      for (int j = 0; j < NT; j++) {
        const int col = j0 + (WN * j + wn) * 8 + tq * 2;
        double2 o = make_double2(0.0, 0.0);
        if (2 * i + wm - 4 * j - wn >= sym_thr) o = *reinterpret_cast<const double2*>(Cg + (long long)row * p.ldc + col);
        acc[i][j][0] = r * o.x;
        acc[i][j][1] = r * o.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
      for (int j = 0; j < NT; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  }

  TG_STAMP(n, 1);
  if (KT > 0) {
    tg_mbar_wait(bar_base + 8 * cs, cph);
    load_frags(0, smem_base + cs * TG_STAGE_BYTES, 0);
  }
  TG_STAMP(n, 2);
  TG_CTA(1);
  // One k-step (16 deep = 4 k4 steps) with the compile-time set of live fragments `Lv`: straight-line DMMAs, no
  // per-instruction predicates (a first version predicated every DMMA: the compiler guards each with WARPSYNC + a
  // chain of ISETPs and the "skipped" work cost more than doing it, profiles/r01_c5_trisk_v1.txt).
  // The four k4 steps of a k-step run as a ROLLED loop of two double-steps (fragment buffers 0 / 1 are compile-time
  // inside an iteration).  Fully unrolled, a body was ~5 KB of SASS and the kernel 150 KB: the eight role bodies of a
  // symmetric diagonal tile (one per warp) or the eight bodies a triangular K block walks through (one per k-step,
  // each executed once per tile) exceed the 32 KB L1.5 instruction cache - `ncu --set full` of the batched
  // S = M^T M launch: 20 % of the warp samples stalled on `no_instructions`, a symmetric 128-deep block took 28k clk
  // for 17.4k clk of DMMA issue (profiles/r02_lauum_icache.txt).  Halving every body brings both working sets under it.
  auto kt_body = [&](auto lv, int kt) {
    using Lv = decltype(lv);
    const int s = cs;
    const uint32_t stage_base = smem_base + s * TG_STAGE_BYTES;
    const int s2 = (cs + 1 == TG_STAGES) ? 0 : cs + 1;
    const uint32_t ph2 = (cs + 1 == TG_STAGES) ? cph ^ 1u : cph;
    const bool has_next = kt + 1 < KT;
    uint32_t rel = 0;
    auto mma = [&](const int buf) {
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++)
          if (Lv::live(i, j)) dmma884(acc[i][j][0], acc[i][j][1], af[buf][i], bf[buf][j]);
    };
    auto fold = [&](const int buf) {
#pragma unroll
      for (int i = 0; i < MT; i++) rel |= (uint32_t)__double2loint(af[buf][i]);
#pragma unroll
      for (int j = 0; j < NT; j++) rel |= (uint32_t)__double2loint(bf[buf][j]);
    };
#pragma unroll 1
    for (int kh = 0; kh < TG_BK / 8; kh++) {
      // k4 step 2 kh: fragments in buffer 0, step 2 kh + 1 requested into buffer 1
      load_frags(1, stage_base, 2 * kh + 1);
      mma(0);
#if !TG_RELEASE_LATE
      fold(0);
      if (kh == TG_BK / 8 - 1) {
        fold(1);
        __syncwarp();
        if (lane == 0) tg_mbar_arrive(bar_base + 64 + 8 * s + (rel & (uint32_t)p.zero));
      }
#endif
      // k4 step 2 kh + 1: buffer 1; the step after it (same stage, or the first of the next stage) goes into buffer 0
      if (kh < TG_BK / 8 - 1) {
        load_frags(0, stage_base, 2 * kh + 2);
      } else if (has_next) {
        tg_mbar_wait(bar_base + 8 * s2, ph2);
        load_frags(0, smem_base + s2 * TG_STAGE_BYTES, 0);
      }
      mma(1);
#if !TG_RELEASE_LATE
      fold(1);
#endif
    }
#if TG_RELEASE_LATE
    // Stage release AFTER the DMMAs of the stage's last k4 step (program order between the asm volatile statements is
    // kept by the compiler, a warp issues in order, and a DMMA cannot issue before the ld.shared that feed its operands
    // have returned): every fragment load of this stage has completed when the arrive issues, so the TMA refill
    // (async-proxy write) cannot overtake a generic-proxy read of the stage.  No data-dependency trick needed.
    (void)rel; (void)fold;
    __syncwarp();
    if (lane == 0) tg_mbar_arrive(bar_base + 64 + 8 * s);
#endif
    cs = s2;
    cph = ph2;
  };
  if (!sym_diag && tri_mode == 0 && !staggered) {
    // dense tile (every long-K tile of the single large factorisation): no per-k-step variant dispatch
#pragma unroll 1
    for (int kt = 0; kt < KT; kt++) kt_body(TgLive<0, MT, 0, NT, -1>{}, kt);
  } else
  for (int kt = 0; kt < KT; kt++) {
#ifdef FFGP_TG_TRACE
    if (n == 0 && kt < 8) TG_STAMP(24 + kt, 0);          // first tile: start of every k-step (rows 24.. of the trace)
#endif
    if (staggered && n == 0 && warp == 0 && lane == 0 && kt == min(p.stagger, KT - 1)) *stagger_go = 1;
    // variant of this k-step (warp-uniform): 0 dense | 1..7 i >= v | 8..14 i < v-7 | 15..17 j >= v-14 | 18..20 j < v-17 |
    // 21..28 symmetric diagonal tile, by warp role.  Ranges are the union over the four k4 steps of the k-step.
    int variant = 0;
    if (sym_diag) {
      variant = 21 + role;
    } else if (tri_mode != 0 && kt >= tri_kt0 && kt < tri_kt0 + ((tri_mode == 1 || tri_mode == 4) ? TG_BM : TG_BN) / TG_BK) {
      const int pk = (kt - tri_kt0) * TG_BK;             // first k of the step inside the diagonal K block
      if (tri_mode == 1) {                               // (2 i + wm) * 8 + 7 >= pk
        const int lo = min(MT - 1, (pk - 7 - 8 * wm + 15) >> 4);
        if (lo > 0) variant = lo;
      } else if (tri_mode == 4) {                        // (2 i + wm) * 8 <= pk + 15
        const int hi = max(1, ((pk + 15 - 8 * wm) >> 4) + 1);
        if (hi < MT) variant = 7 + hi;
      } else if (tri_mode == 2) {                        // (4 j + wn) * 8 + 7 >= pk
        const int lo = min(NT - 1, (pk - 7 - 8 * wn + 8 * WN - 1) / (8 * WN));
        if (lo > 0) variant = 14 + lo;
      } else {                                           // (4 j + wn) * 8 <= pk + 15
        const int hi = max(1, (pk + 15 - 8 * wn) / (8 * WN) + 1);
        if (hi < NT) variant = 17 + hi;
      }
    }
    switch (variant) {
      case 0: kt_body(TgLive<0, MT, 0, NT, -1>{}, kt); break;
      case 1: kt_body(TgLive<1, MT, 0, NT, -1>{}, kt); break;
      case 2: kt_body(TgLive<2, MT, 0, NT, -1>{}, kt); break;
      case 3: kt_body(TgLive<3, MT, 0, NT, -1>{}, kt); break;
      case 4: kt_body(TgLive<4, MT, 0, NT, -1>{}, kt); break;
      case 5: kt_body(TgLive<5, MT, 0, NT, -1>{}, kt); break;
      case 6: kt_body(TgLive<6, MT, 0, NT, -1>{}, kt); break;
      case 7: kt_body(TgLive<7, MT, 0, NT, -1>{}, kt); break;
      case 8: kt_body(TgLive<0, 1, 0, NT, -1>{}, kt); break;
      case 9: kt_body(TgLive<0, 2, 0, NT, -1>{}, kt); break;
      case 10: kt_body(TgLive<0, 3, 0, NT, -1>{}, kt); break;
      case 11: kt_body(TgLive<0, 4, 0, NT, -1>{}, kt); break;
      case 12: kt_body(TgLive<0, 5, 0, NT, -1>{}, kt); break;
      case 13: kt_body(TgLive<0, 6, 0, NT, -1>{}, kt); break;
      case 14: kt_body(TgLive<0, 7, 0, NT, -1>{}, kt); break;
      case 15: kt_body(TgLive<0, MT, 1, NT, -1>{}, kt); break;
      case 16: kt_body(TgLive<0, MT, 2, NT, -1>{}, kt); break;
      case 17: kt_body(TgLive<0, MT, 3, NT, -1>{}, kt); break;
      case 18: kt_body(TgLive<0, MT, 0, 1, -1>{}, kt); break;
      case 19: kt_body(TgLive<0, MT, 0, 2, -1>{}, kt); break;
      case 20: kt_body(TgLive<0, MT, 0, 3, -1>{}, kt); break;
      case 21: kt_body(TgLive<0, MT, 0, NT, 0>{}, kt); break;
      case 22: kt_body(TgLive<0, MT, 0, NT, 1>{}, kt); break;
      case 23: kt_body(TgLive<0, MT, 0, NT, 2>{}, kt); break;
      case 24: kt_body(TgLive<0, MT, 0, NT, 3>{}, kt); break;
      case 25: kt_body(TgLive<0, MT, 0, NT, 4>{}, kt); break;
      case 26: kt_body(TgLive<0, MT, 0, NT, 5>{}, kt); break;
      case 27: kt_body(TgLive<0, MT, 0, NT, 6>{}, kt); break;
      default: kt_body(TgLive<0, MT, 0, NT, 7>{}, kt); break;
    }
  }

  if (staggered && n == 0 && KT == 0 && warp == 0 && lane == 0) *stagger_go = 1;
  TG_STAMP(n, 3);
  TG_CTA(2);
  // ---- epilogue: C fragment (row g, cols 2 tq, 2 tq + 1) -> 16-byte stores ----------------------
#pragma unroll
  for (int i = 0; i < MT; i++) {
    const int row = i0 + (2 * i + wm) * 8 + g;
#pragma unroll
    for (int j = 0; j < NT; j++) {
      if (2 * i + wm - 4 * j - wn < sym_thr) continue;      // above the diagonal of a symmetric tile: not produced
      const int col = j0 + (WN * j + wn) * 8 + tq * 2;
      double2* dst = reinterpret_cast<double2*>(Cg + (long long)row * p.ldc + col);
      double2 v = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
      if (beta != 0.0 && !init_from_c) {        // alpha == 0: C is only scaled
        const double2 o = *dst;
        v.x = fma(beta, o.x, v.x);
        v.y = fma(beta, o.y, v.y);
      }
      *dst = v;
    }
  }
  TG_STAMP(n, 4);
  TG_CTA(3);
  if (!p.sched) break;
  // next work item of this CTA: its id is valid once the first stage of the tile (or the end marker) has been posted
  tg_mbar_wait(bar_base + 8 * cs, cph);
  w = tile_ring[(n + 1) % TG_TILE_RING];
  TG_STAMP(n, 5);
  if (w < 0) break;
  }   // tile loop
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor-map encoding (driver entry point fetched through the runtime, no libcuda link) and launch.
// ---------------------------------------------------------------------------------------------
typedef CUresult (*TgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TgEncodeFn tg_encode_fn() {
  static TgEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (TgEncodeFn)f;
  }
  return fn;
}

// rows x K operand with row/col layout flag, leading dimension ld, two batch levels (inner count/stride, outer
// count/stride; strides in elements).  Returns false when the operand cannot be described (caller falls back).
inline bool tg_make_map(CUtensorMap* map, const double* base, bool kmaj, int rows, int K, int ld, int inner,
                        long long istride, int outer, long long ostride, int box_rows = TG_BM) {
  TgEncodeFn enc = tg_encode_fn();
  if (!enc) return false;
  if (((uintptr_t)base & 15) || (ld & 1)) return false;
  if ((inner > 1 && (istride <= 0 || (istride & 1))) || (outer > 1 && (ostride <= 0 || (ostride & 1)))) return false;
  const cuuint64_t ib = (inner > 1 ? (cuuint64_t)istride : 2) * 8, ob = (outer > 1 ? (cuuint64_t)ostride : 2) * 8;
  CUresult r;
  if (kmaj) {
    cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 8, ib, ob};
    cuuint32_t box[4] = {(cuuint32_t)TG_BK, (cuuint32_t)box_rows, 1, 1}, es[4] = {1, 1, 1, 1};
    r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {8, (cuuint64_t)K, (cuuint64_t)(rows / 8), (cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[4] = {(cuuint64_t)ld * 8, 64, ib, ob};
    cuuint32_t box[5] = {8, (cuuint32_t)TG_BK, (cuuint32_t)(box_rows / 8), 1, 1}, es[5] = {1, 1, 1, 1, 1};
    r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS;
}

template <bool A_KMAJ, bool B_KMAJ, int WN = 4>
cudaError_t launch_gemm_tma_cfg(const CUtensorMap& mA, const CUtensorMap& mB, const TmaGemmParams& tp, int grid,
                                cudaStream_t st) {
  auto kern = gemm_tma_kernel<A_KMAJ, B_KMAJ, WN>;
  constexpr size_t TG_SMEM_BYTES = TgCfg<WN>::SMEM_BYTES;
  constexpr int TG_THREADS = TgCfg<WN>::THREADS;
  static PerDeviceOnce once;
  bool& attr_set = *once.slot();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  kern<<<grid, TG_THREADS, TG_SMEM_BYTES, st>>>(mA, mB, tp);
  return cudaGetLastError();
}

// Returns cudaErrorNotSupported when the problem cannot go through the TMA kernel (caller uses gemm_dmma_kernel).
// Work-item counters of the persistent grids: ONE {next, done} pair PER STREAM, zero at load and rearmed by the last CTA
// of each launch.  Kernels of one stream run one after the other, so a pair is never shared by two launches in flight;
// launches on different streams (the library's own side streams, several host threads, a graph replay next to eager
// work) get different pairs.  Under stream capture the persistent mode is switched off: a captured node would freeze
// the capture stream's pair into the graph, and the graph may later be replayed on any stream.
constexpr int TG_SCHED_SLOTS = 64;
__device__ unsigned int g_tg_sched[2 * TG_SCHED_SLOTS];
struct TgSchedTable {
  std::mutex mu;
  unsigned int* base[64] = {nullptr};
  cudaStream_t streams[64][TG_SCHED_SLOTS] = {};
  int used[64] = {};
  // counter pair of (current device, stream); nullptr: table full or symbol unavailable (caller launches non-persistent)
  unsigned int* slot(cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    dev &= 63;
    std::lock_guard<std::mutex> lock(mu);
    if (!base[dev] && cudaGetSymbolAddress((void**)&base[dev], g_tg_sched) != cudaSuccess) { base[dev] = nullptr; return nullptr; }
    for (int i = 0; i < used[dev]; i++)
      if (streams[dev][i] == st) return base[dev] + 2 * i;
    if (used[dev] >= TG_SCHED_SLOTS) return nullptr;
    streams[dev][used[dev]] = st;
    return base[dev] + 2 * used[dev]++;
  }
};
inline TgSchedTable& tg_sched_table() {
  static TgSchedTable t;
  return t;
}

// `persistent_ctas` > 0: that many CTAs pull work items from a launch-wide counter, so the TMA producer streams the next
// tile's panels while the MMA warps finish and store the current one (no per-tile launch / barrier-init / first-load
// latency; batched small problems, where a tile is only 64-256 deep).  0: one CTA per work item, heavy tiles first -
// the hardware scheduler balances and SMs free up for a high-priority panel chain as tiles retire.
// `narrow`: 128 x 64 tiles, two CTAs per SM (TgCfg<2>); needs N % 64 == 0 and no lower_only.
inline cudaError_t launch_gemm_tma(bool a_kmaj, bool b_kmaj, const GemmParams& p, int batch_outer, cudaStream_t st,
                                   int persistent_ctas = 0, bool narrow = false) {
  const int BN = narrow ? 64 : TG_BN;
  if (narrow && p.lower_only) return cudaErrorNotSupported;
  if (p.M % TG_BM || p.N % BN || p.K % TG_BK || p.M <= 0 || p.N <= 0 || p.K <= 0) return cudaErrorNotSupported;
  if (((uintptr_t)p.C & 15) || (p.ldc & 1)) return cudaErrorNotSupported;
  CUtensorMap mA, mB;
  if (!tg_make_map(&mA, p.A, a_kmaj, p.M, p.K, p.lda, p.inner, p.iA, batch_outer, p.sA)) return cudaErrorNotSupported;
  if (!tg_make_map(&mB, p.B, b_kmaj, p.N, p.K, p.ldb, p.inner, p.iB, batch_outer, p.sB, BN)) return cudaErrorNotSupported;
  TmaGemmParams tp;
  tp.C = p.C; tp.ldc = p.ldc; tp.sC = p.sC; tp.iC = p.iC; tp.M = p.M; tp.N = p.N; tp.K = p.K; tp.inner = p.inner;
  tp.alpha = p.alpha; tp.beta = p.beta; tp.lower_only = p.lower_only; tp.kmode = p.kmode; tp.heavy_first = p.heavy_first;
  const int tm = p.M / TG_BM, tn = p.N / BN;
  const int tiles = p.lower_only ? tm * (tm + 1) / 2 : tm * tn;
  const long long total = (long long)tiles * batch_outer * p.inner;
  if (total > 0x7fffffffLL) return cudaErrorNotSupported;
  tp.tiles = tiles; tp.total = (int)total; tp.zero = 0;
  {
    static int stg = -1;                                 // FFGP_STAGGER: k-steps of offset between the two warp rows (0 = off)
    if (stg < 0) { const char* e = getenv("FFGP_STAGGER"); stg = e ? atoi(e) : 0; if (stg < 0 || stg >= TG_STAGES) stg = 0; }
    tp.stagger = narrow ? 0 : stg;
  }
  int grid = (int)total;
  tp.sched = nullptr;
  if (narrow) persistent_ctas *= 2;                      // two CTAs per SM
  if (persistent_ctas > 0 && total > persistent_ctas) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
      tp.sched = tg_sched_table().slot(st);
      if (tp.sched) grid = persistent_ctas;
    }
  }
  if (narrow) {
    if (a_kmaj && b_kmaj) return launch_gemm_tma_cfg<true, true, 2>(mA, mB, tp, grid, st);
    if (a_kmaj && !b_kmaj) return launch_gemm_tma_cfg<true, false, 2>(mA, mB, tp, grid, st);
    if (!a_kmaj && b_kmaj) return launch_gemm_tma_cfg<false, true, 2>(mA, mB, tp, grid, st);
    return launch_gemm_tma_cfg<false, false, 2>(mA, mB, tp, grid, st);
  }
  if (a_kmaj && b_kmaj) return launch_gemm_tma_cfg<true, true>(mA, mB, tp, grid, st);
  if (a_kmaj && !b_kmaj) return launch_gemm_tma_cfg<true, false>(mA, mB, tp, grid, st);
  if (!a_kmaj && b_kmaj) return launch_gemm_tma_cfg<false, true>(mA, mB, tp, grid, st);
  return launch_gemm_tma_cfg<false, false>(mA, mB, tp, grid, st);
}

}  // namespace ffgp
