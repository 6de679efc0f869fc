// Host side of the dense GP path: workspace layout, the recursive potrf+trtri driver and the
// C-ABI entry points declared in include/ffgp.h.  No allocation, no host synchronisation.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../include/ffgp.h"
#include "dense_kernels.cuh"
#include "gemm_dmma.cuh"
#include "gemm_tma.cuh"
#include "factor256.cuh"
#include "xgrad_kernels.cuh"
#include "acq_kernels.cuh"
#include "train_kernels.cuh"
#include "match_kernels.cuh"
#include "pack_kernels.cuh"

namespace ffgp {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};   // kernels launched by this library (bench.py reports it as gpu_launches)
// FFGP_TRACE=1: a timing event after every level-3 launch of a dense evaluation, dumped (with a host sync) by
// ffgp_trace_dump() - the timeline tool behind profiles/r01_timeline_*.txt.  Never enabled in production.
struct TraceRec { const char* what; int a, b; void* stream; cudaEvent_t ev; };
static TraceRec g_trace[8192];
static int g_ntrace = 0;
static int trace_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_TRACE"); v = e ? atoi(e) : 0; }
  return v;
}
static void trace_mark(const char* what, int a, int b, cudaStream_t st) {
  if (!trace_on() || g_ntrace >= 8192) return;
  TraceRec& r = g_trace[g_ntrace++];
  r.what = what; r.a = a; r.b = b; r.stream = (void*)st;
  cudaEventCreate(&r.ev);
  cudaEventRecord(r.ev, st);
}

int fail(int code, const char* fmt, const char* a = "") {
  snprintf(g_err, sizeof(g_err), fmt, a);
  return code;
}
#define FFGP_CUDA(x)                                                   \
  do {                                                                 \
    cudaError_t e__ = (x);                                             \
    if (e__ != cudaSuccess) return fail(-100, "CUDA error: %s", cudaGetErrorString(e__)); \
  } while (0)

#define FFGP_LAUNCHED()          \
  do {                           \
    ++ffgp::g_launches;          \
    FFGP_CUDA(cudaGetLastError()); \
  } while (0)

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

// ---------------------------------------------------------------------------------------------
// GEMM dispatch: layout flags -> template instance; tile size by divisibility and machine fill
// ---------------------------------------------------------------------------------------------
static int num_sms() {
  static PerDeviceInt per_dev;
  int& v = *per_dev.slot();
  if (v == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

// FFGP_GEMM_TMA=0 routes the 128-tile GEMMs through the older cp.async kernel (A/B comparisons only)
static bool use_tma_gemm() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_GEMM_TMA"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}

static thread_local int g_cta_cap = 0;   // > 0: see gemm()
// Batched small problems (no priority chain to protect): the TMA GEMM runs as a persistent grid, see launch_gemm_tma.
// Measured per launch at 512 x (N = 512) (profiles/r01_launches_c5_v7_persistent.txt): 5-9 % faster for the products
// with beta = 0, but slower where the epilogue has to read C (SYRK updates: +25 %) and for the lower-tiles-only
// M^T M product (+20 %), so only the former use it.
// FFGP_PERSIST=0 disables it (A/B comparisons), FFGP_PERSIST=2 forces it for every launch (tests/test_gpu_gemm.py).
static thread_local int g_persist = 0;
// > 0: the launch runs as a persistent grid on all SMs but this many, which stay free for the high-priority panel chain
// of a single large factorisation (see potrf_right_looking).  FFGP_SYRK_RESERVE sets it.
static thread_local int g_reserve_sms = 0;
static int persist_lower() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_PERSIST_LOWER"); v = e ? atoi(e) : 0; }
  return v;
}
static int sched_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_SCHED"); v = e ? atoi(e) : 1; }
  return v;
}
static int split_colupd() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_SPLIT_COLUPD"); v = e ? atoi(e) : 0; }
  return v;
}
// FFGP_BASE64=1: batched problems bottom out at 64-blocks (experiment, see factor_rec)
static int batched_base64() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_BASE64"); v = e ? atoi(e) : 0; }
  return v;
}
static int syrk_reserve() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_SYRK_RESERVE"); v = e ? atoi(e) : 0; }
  return v;
}
static int persist_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_PERSIST"); v = e ? atoi(e) : 1; }
  return v;
}
// FFGP_F256: 1 (default) batched problems factor and invert a 256 x 256 diagonal block with ONE launch of
// factor256_kernel (csrc/factor256.cuh: one CTA per problem runs the two base blocks and the four 128-cube products
// between them without leaving the SM); 2: the same without the structural-zero skipping; 3: also for single problems;
// 0: two base kernels + four GEMM launches.  Measured (profiles/r02_factor256.txt): 592 x (N = 512) potrf + trtri
// 3.53 -> 3.41 ms; a single problem does NOT gain - its four products spread over several SMs as 64-tiles and the chain
// of a right-looking step at N = 8192 is 150 us either way (20.33 vs 20.24 ms per evaluation) - so it keeps the six launches.
static int f256_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_F256"); v = e ? atoi(e) : 1; }
  return v;
}
static thread_local const char* g_trace_label = "gemm";

static cudaError_t gemm(bool a_kmaj, bool b_kmaj, const double* A, int lda, long long sA, const double* B, int ldb,
                        long long sB, double* C, int ldc, long long sC, int M, int N, int K, double alpha, double beta,
                        int lower_only, int kmode, int batch, cudaStream_t st, int inner = 1, long long iA = 0,
                        long long iB = 0, long long iC = 0) {
  // Background launches (g_cta_cap > 0) are cut into bands of at most g_cta_cap CTAs: a grid that fills every SM with
  // long-K tiles would make the high-priority panel chain wait for a tile to retire at every one of its launches
  // (a 128x128 TMA GEMM CTA owns its SM's whole register file).  Bands run back to back on the same stream.
  if (g_cta_cap > 0 && !lower_only && batch == 1 && M % 128 == 0 && N % 128 == 0) {
    const int cap = g_cta_cap;
    const int tm = M / 128, tn = N / 128;
    if ((long long)tm * tn * inner > cap) {
      g_cta_cap = 0;
      cudaError_t e = cudaSuccess;
      if (inner > 1) {
        const int per = std::max(1, cap / (tm * tn));
        for (int n0 = 0; n0 < inner && e == cudaSuccess; n0 += per) {
          g_cta_cap = (tm * tn > cap) ? cap : 0;       // a single node larger than the cap is banded by the recursion
          e = gemm(a_kmaj, b_kmaj, A + n0 * iA, lda, sA, B + n0 * iB, ldb, sB, C + n0 * iC, ldc, sC, M, N, K, alpha, beta,
                   lower_only, kmode, 1, st, std::min(per, inner - n0), iA, iB, iC);
          g_cta_cap = 0;
        }
      } else if (kmode == K_LE_ROW || kmode == K_GE_ROW) {   // K range depends on the tile ROW: band along N
        const int per = std::max(1, cap / tm) * 128;
        for (int c0 = 0; c0 < N && e == cudaSuccess; c0 += per)
          e = gemm(a_kmaj, b_kmaj, A, lda, sA, b_kmaj ? B + (long long)c0 * ldb : B + c0, ldb, sB, C + c0, ldc, sC, M,
                   std::min(per, N - c0), K, alpha, beta, 0, kmode, 1, st);
      } else {                                               // K range depends on the tile COLUMN (or not at all): band along M
        const int per = std::max(1, cap / tn) * 128;
        for (int r0 = 0; r0 < M && e == cudaSuccess; r0 += per)
          e = gemm(a_kmaj, b_kmaj, a_kmaj ? A + (long long)r0 * lda : A + r0, lda, sA, B, ldb, sB, C + (long long)r0 * ldc, ldc,
                   sC, std::min(per, M - r0), N, K, alpha, beta, 0, kmode, 1, st);
      }
      g_cta_cap = cap;
      return e;
    }
  }
  GemmParams p;
  p.A = A; p.B = B; p.C = C; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.sA = sA; p.sB = sB; p.sC = sC; p.alpha = alpha; p.beta = beta; p.lower_only = lower_only; p.kmode = kmode;
  p.heavy_first = 1;
  p.inner = inner; p.iA = iA; p.iB = iB; p.iC = iC;
  const int batch_outer = batch;
  batch *= inner;
  ++g_launches;
  bool big = (M % 128 == 0) && (N % 128 == 0);
  if (big) {
    const long long tm = M / 128, tn = N / 128;
    const long long tiles = (lower_only ? tm * (tm + 1) / 2 : tm * tn) * batch;
    if (tiles < (long long)num_sms()) big = false;     // 64x64 tiles: 4x the CTAs for the small levels
  }
  cudaError_t e = cudaErrorNotSupported;
  int persistent_ctas = (persist_mode() == 2 || (g_persist && persist_mode() == 1 && (!lower_only || persist_lower()))) ? num_sms() : 0;
  if (g_reserve_sms > 0 && g_reserve_sms < num_sms()) persistent_ctas = num_sms() - g_reserve_sms;
  // 128 x 64 tiles with two CTAs per SM (gemm_tma.cuh TgCfg<2>): the short-K products of batched problems (a tile's
  // hand-over is 15-25 % of a 128/256-deep tile and one CTA per SM cannot hide it) and 64-wide right-hand sides.
  // FFGP_GEMM_N64: 0 never, 1 (default) batched non-lower launches with K <= 1024 and N % 128 != 0 shapes, 2 every eligible launch.
  static int n64 = -1;
  if (n64 < 0) { const char* ev = getenv("FFGP_GEMM_N64"); n64 = ev ? atoi(ev) : 1; }
  const bool narrow_ok = use_tma_gemm() && !lower_only && M % 128 == 0 && N % 64 == 0 && K % 16 == 0 &&
                         (long long)(M / 128) * (N / 64) * batch >= 2LL * num_sms();
  const bool narrow = narrow_ok && (n64 == 2 || (n64 == 1 && ((g_persist && K <= 1024) || N % 128 != 0)));
  if (narrow) e = launch_gemm_tma(a_kmaj, b_kmaj, p, batch_outer, st, persistent_ctas, true);
  if (e == cudaErrorNotSupported && big && use_tma_gemm()) e = launch_gemm_tma(a_kmaj, b_kmaj, p, batch_outer, st, persistent_ctas);
  if (e == cudaErrorNotSupported) {
    if (a_kmaj && b_kmaj) e = launch_gemm<true, true>(p, batch, big, st);
    else if (a_kmaj && !b_kmaj) e = launch_gemm<true, false>(p, batch, big, st);
    else if (!a_kmaj && b_kmaj) e = launch_gemm<false, true>(p, batch, big, st);
    else e = launch_gemm<false, false>(p, batch, big, st);
  }
  trace_mark(g_trace_label, M, N * 100000 + K, st);
  return e;
}

// ---------------------------------------------------------------------------------------------
// Workspace
// ---------------------------------------------------------------------------------------------
struct DenseWs {
  int np, Dp, nsp, chunk, nblk, ngtile;
  bool gemm_rhs;          // D > 8: right-hand sides go through the GEMM kernel
  double *A, *L, *M;      // [chunk][np][np]
  double *Gm, *alpha;     // [chunk][np][Dw]   (Dw = D when !gemm_rhs, else Dp)
  double* Ypad;           // [chunk][np][Dp]   (gemm_rhs only)
  double *rowsq;          // [chunk][np]
  double *logdet_part;    // [chunk][nblk]
  double *partial;        // [chunk][ngtile][d+1]
  double *Kx, *V;         // [chunk][np][nsp]
  double *Kxx;            // [chunk][nsp][nsp]
  double *colsq;          // [chunk][nsp]
  double *meanp;          // [chunk][nsp][Dp]
  size_t bytes;
};

// Per-chunk working set bound for big batches (FFGP_WS_GB, default 32 GiB of the B200's 180 GB).  Every launch of a chunk
// streams the whole chunk through HBM whatever its size, so a smaller chunk buys no locality; it only multiplies the ~35
// dependent launches of an evaluation (each with ~11 us of ramp-up and drain) by the number of chunks.  BASELINE config 5
// (4096 x N = 512): 5 chunks of <= 888 problems under the round-1 bound of 6 GiB, ONE chunk of 26 GiB now.
static size_t ws_target_bytes() {
  static size_t v = 0;
  if (!v) {
    const char* e = getenv("FFGP_WS_GB");
    const long long gb = e ? atoll(e) : 32;
    v = (size_t)(gb < 1 ? 1 : gb) << 30;
  }
  return v;
}

static DenseWs layout_ws(int n, int d, int D, int ns, int batch, char* base) {
  DenseWs w;
  w.np = round_up(std::max(n, 1), 128);
  w.gemm_rhs = D > 8;
  w.Dp = round_up(std::max(D, 1), 64);
  w.nsp = ns > 0 ? round_up(ns, 64) : 0;
  w.nblk = w.np / BASE_N_BATCHED;
  const int gt = w.np / GRAD_T;
  w.ngtile = gt * (gt + 1) / 2;
  const int Dw = w.gemm_rhs ? w.Dp : D;
  auto per_item = [&]() {
    size_t s = 0;
    s += 3 * align256((size_t)w.np * w.np * 8);
    s += 2 * align256((size_t)w.np * Dw * 8);
    if (w.gemm_rhs) s += align256((size_t)w.np * w.Dp * 8);
    s += align256((size_t)w.np * 8) + align256((size_t)w.nblk * 8) + align256((size_t)w.ngtile * (d + 1) * 8);
    if (w.nsp) {
      s += 2 * align256((size_t)w.np * w.nsp * 8) + align256((size_t)w.nsp * w.nsp * 8) + align256((size_t)w.nsp * 8);
      s += align256((size_t)w.nsp * w.Dp * 8);
    }
    return s;
  };
  const size_t item = per_item();
  long long chunk = (long long)(ws_target_bytes() / item);
  if (chunk < 1) chunk = 1;
  {  // balanced chunks: 1024 problems with room for 1016 must not become 1016 + 8 (the tail chunk pays every launch
     // latency of a full one, profiles/r01_metrics_c5_v1.txt)
    const long long bt = std::max(batch, 1);
    const long long nchunks = (bt + chunk - 1) / chunk;
    w.chunk = (int)((bt + nchunks - 1) / nchunks);
    // ... and whole waves: a third of the launches of a batched evaluation have ONE CTA per problem (base blocks, 128-level
    // products), so a chunk of 820 problems is 5.54 waves on 148 SMs = 6 rounds.  Prefer the largest multiple of the SM
    // count that fits when it does not add a chunk (4096 problems: 4 x 888 + 544 = 28 rounds instead of 5 x 820 = 30).
    const long long sms = 148;
    const long long whole = chunk / sms * sms;
    if (nchunks > 1 && whole >= sms && (bt + whole - 1) / whole <= nchunks) w.chunk = (int)whole;
  }
  size_t off = 0;
  auto take = [&](size_t nbytes) {
    char* ptr = base ? base + off : nullptr;
    off += align256(nbytes) ;
    return reinterpret_cast<double*>(ptr);
  };
  const size_t c = (size_t)w.chunk;
  w.A = take(c * w.np * w.np * 8);
  w.L = take(c * w.np * w.np * 8);
  w.M = take(c * w.np * w.np * 8);
  w.Gm = take(c * w.np * Dw * 8);
  w.alpha = take(c * w.np * Dw * 8);
  w.Ypad = w.gemm_rhs ? take(c * w.np * w.Dp * 8) : nullptr;
  w.rowsq = take(c * w.np * 8);
  w.logdet_part = take(c * w.nblk * 8);
  w.partial = take(c * w.ngtile * (d + 1) * 8);
  if (w.nsp) {
    w.Kx = take(c * w.np * w.nsp * 8);
    w.V = take(c * w.np * w.nsp * 8);
    w.Kxx = take(c * w.nsp * w.nsp * 8);
    w.colsq = take(c * w.nsp * 8);
    w.meanp = take(c * w.nsp * w.Dp * 8);
  } else {
    w.Kx = w.V = w.Kxx = w.colsq = w.meanp = nullptr;
  }
  w.bytes = off + 256;
  return w;
}

// The same layout advanced by k problems (every array is [chunk][...]): the slice a second stream works on.
static DenseWs ws_offset(const DenseWs& w, int k, int d, int D) {
  DenseWs r = w;
  const size_t np = w.np, nsp = w.nsp, Dw = w.gemm_rhs ? w.Dp : D, kk = (size_t)k;
  r.A += kk * np * np; r.L += kk * np * np; r.M += kk * np * np;
  r.Gm += kk * np * Dw; r.alpha += kk * np * Dw;
  if (r.Ypad) r.Ypad += kk * np * w.Dp;
  r.rowsq += kk * np; r.logdet_part += kk * w.nblk; r.partial += kk * w.ngtile * (d + 1);
  if (w.nsp) { r.Kx += kk * np * nsp; r.V += kk * np * nsp; r.Kxx += kk * nsp * nsp; r.colsq += kk * nsp; r.meanp += kk * nsp * w.Dp; }
  return r;
}

// ---------------------------------------------------------------------------------------------
// Recursive blocked factorisation:  (L, M = L^-1) of the block [off, off+n) of A.
//   F(A11) ; L21 = A21 M11^T ; A22 -= L21 L21^T ; F(A22) ; T = M22 L21 (into the dead A21) ; M21 = -T M11
// Every step above the 128x128 base is one DMMA GEMM launch over the whole batch; the triangular
// operands only shorten the K range of a tile, never add a special kernel.
// ---------------------------------------------------------------------------------------------
struct FactorCtx {
  double *A, *L, *M;
  int ld;
  long long sb;           // batch stride (elements)
  int batch;
  double* logdet_part; int nblk;   // partial log-determinants, one slot per BASE_N_BATCHED rows
  int* info;
  cudaStream_t st;
  int base_n = BASE_N;             // block handled by the base kernel: BASE_N, or BASE_N_BATCHED for batched problems
};

static cudaError_t factor_rec(const FactorCtx& c, int off, int n) {
  cudaError_t e;
  const long long d0 = (long long)off * c.ld + off;
  if (n == c.base_n) {
    // FFGP_BASE64=1 (experiment): batched problems bottom out at 64 instead of 128 - 35 KB of shared memory, two CTAs
    // per SM, the extra level of 64-wide products through the GEMM kernel.  Measured on 512 x (N = 512)
    // (profiles/r01_launches_c5_v11_base64.txt): the base kernels drop from 0.76 to 0.45 ms per chunk, but a 64-block
    // still costs 8 panels x ~3.5 us (the dependent pivot chain and the barriers, not the flops, set the panel time)
    // and the 16 extra GEMM launches add 0.29 ms: no net gain, so the default stays 128.
    if (c.base_n == BASE_N)
      potrf_trtri_base_kernel<BASE_N><<<c.batch, 256, base_smem_bytes(BASE_N), c.st>>>(
          c.A + d0, c.L + d0, c.M + d0, c.ld, c.sb, c.logdet_part, c.nblk, off / BASE_N_BATCHED, c.info, off);
    else
      potrf_trtri_base_kernel<BASE_N_BATCHED><<<c.batch, 256, base_smem_bytes(BASE_N_BATCHED), c.st>>>(
          c.A + d0, c.L + d0, c.M + d0, c.ld, c.sb, c.logdet_part, c.nblk, off / BASE_N_BATCHED, c.info, off);
    ++g_launches;
    trace_mark("base", off, 0, c.st);
    return cudaGetLastError();
  }
  if (n == 2 * BASE_N && c.base_n == BASE_N && f256_mode() != 0 && (c.batch >= 8 || f256_mode() == 3)) {
    factor256_kernel<<<c.batch, 256, F2_SMEM, c.st>>>(c.A + d0, c.L + d0, c.M + d0, c.ld, c.sb, c.logdet_part, c.nblk,
                                                      off / BASE_N_BATCHED, c.info, off, f256_mode() == 2 ? 0 : 1);
    ++g_launches;
    trace_mark("f256", off, 0, c.st);
    return cudaGetLastError();
  }
  // split at a multiple of the base block: top half gets the larger power-of-two-ish share
  int h = ((n / c.base_n) + 1) / 2 * c.base_n;
  const int m = n - h;
  if ((e = factor_rec(c, off, h)) != cudaSuccess) return e;
  const long long o21 = (long long)(off + h) * c.ld + off;
  const long long o22 = (long long)(off + h) * c.ld + off + h;
  // L21 = A21 * M11^T          (M11[j][p] = 0 for p > j)
  if ((e = gemm(true, true, c.A + o21, c.ld, c.sb, c.M + d0, c.ld, c.sb, c.L + o21, c.ld, c.sb, m, h, h, 1.0, 0.0, 0,
                K_LE_COL, c.batch, c.st)) != cudaSuccess) return e;
  // A22 -= L21 L21^T           (lower tiles)
  if ((e = gemm(true, true, c.L + o21, c.ld, c.sb, c.L + o21, c.ld, c.sb, c.A + o22, c.ld, c.sb, m, m, h, -1.0, 1.0, 1,
                K_FULL, c.batch, c.st)) != cudaSuccess) return e;
  if ((e = factor_rec(c, off + h, m)) != cudaSuccess) return e;
  // T = M22 * L21  -> A21      (M22[i][p] = 0 for p > i)
  if ((e = gemm(true, false, c.M + o22, c.ld, c.sb, c.L + o21, c.ld, c.sb, c.A + o21, c.ld, c.sb, m, h, m, 1.0, 0.0, 0,
                K_LE_ROW, c.batch, c.st)) != cudaSuccess) return e;
  // M21 = -T * M11             (M11[p][j] = 0 for p < j)
  if ((e = gemm(true, false, c.A + o21, c.ld, c.sb, c.M + d0, c.ld, c.sb, c.M + o21, c.ld, c.sb, m, h, h, -1.0, 0.0, 0,
                K_GE_COL, c.batch, c.st)) != cudaSuccess) return e;
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// Production schedule (profiles/r01_launches_c2_v1.csv showed the fully recursive form spending 23 % of an
// evaluation in 64 serialised 128-block kernels and 17 % in 242 small GEMM launches):
//   1. right-looking blocked Cholesky with outer block NB: the diagonal block is factored (and inverted) by
//      factor_rec on a high-priority side stream ONE STEP AHEAD of the trailing update (look-ahead): as soon as the
//      next block column has received its rank-NB update, its diagonal block + TRSM run concurrently with the rest
//      of the SYRK, so the serial panel work leaves the critical path while the trailing matrix is large;
//   2. triangular inverse bottom-up: with the diagonal-block inverses known, M21 = -M22 L21 M11 of ALL nodes of one
//      tree level is a single batched launch pair (inner batch level of the GEMM), 2 launches per level;
//   3. (gradient) S = M^T M, one launch.
// ---------------------------------------------------------------------------------------------
struct AuxStream {
  cudaStream_t st = nullptr;        // panel chain (highest priority)
  cudaStream_t st_bulk = nullptr;   // trailing updates of a large single factorisation (middle priority)
  cudaStream_t st_bg = nullptr;     // background work overlapped with the chain-bound tail (lowest priority)
  cudaEvent_t ev_main = nullptr, ev_aux = nullptr, ev_fork = nullptr, ev_join = nullptr, ev_half = nullptr, ev_bg = nullptr, ev_rest = nullptr;
};
static AuxStream g_aux[64];

static cudaError_t get_aux(AuxStream** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  AuxStream& a = g_aux[dev & 63];
  if (!a.st) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);        // lo = least urgent (0), hi = most urgent (negative)
    if ((e = cudaStreamCreateWithPriority(&a.st, cudaStreamNonBlocking, hi)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithPriority(&a.st_bulk, cudaStreamNonBlocking, (lo + hi) / 2)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithPriority(&a.st_bg, cudaStreamNonBlocking, lo)) != cudaSuccess) return e;
    cudaEvent_t* evs[7] = {&a.ev_main, &a.ev_aux, &a.ev_fork, &a.ev_join, &a.ev_half, &a.ev_bg, &a.ev_rest};
    for (auto ev : evs)
      if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
  }
  *out = &a;
  return cudaSuccess;
}

static int g_outer_nb = 256;     // outer block of the right-looking sweep (128 or 256); FFGP_NB overrides
static int outer_nb() {
  static bool init = false;
  if (!init) {
    const char* e = getenv("FFGP_NB");
    if (e) { int v = atoi(e); if (v == 128 || v == 256 || v == 512) g_outer_nb = v; }
    init = true;
  }
  return g_outer_nb;
}

// Panel widths of the look-ahead sweep of ONE large problem (FFGP_NB_EARLY = e, FFGP_NB_LATE = l, in sixteenths of the
// matrix): columns [0, e np/16) in 512-wide panels (a step there is bound by its trailing update, and a 512-deep tile
// reads and writes C half as often per FLOP), columns [l np/16, np) in 128-wide panels (a step there is bound by the
// serial chain colupd -> diagonal block -> TRSM, which is shorter for a narrower panel), NB in between.
// Defaults 0 / 16 = constant NB (profiles/r02_c2_panel_widths.txt).
struct PanelWidths {
  int nb = 256, early_until = 0, late_from = 1 << 30;
  int at(int c0) const { return c0 < early_until ? 512 : (c0 >= late_from ? BASE_N : nb); }
  int min_width() const { return late_from < (1 << 30) ? BASE_N : nb; }
};
static PanelWidths panel_widths(int np, int NB, bool single) {
  PanelWidths w;
  w.nb = NB;
  static int early = -1, late = -1;
  if (early < 0) { const char* e = getenv("FFGP_NB_EARLY"); early = e ? atoi(e) : 0; if (early < 0 || early > 16) early = 0; }
  if (late < 0) { const char* e = getenv("FFGP_NB_LATE"); late = e ? atoi(e) : 16; if (late < 0 || late > 16) late = 16; }
  if (!single || NB != 256 || np % 512 != 0 || np < 2048) return w;
  const int eu = (int)((long long)np * early / 16), lf = (int)((long long)np * late / 16);
  if (early > 0 && eu % 512 == 0 && eu <= lf) w.early_until = eu;
  if (late < 16 && lf % NB == 0) w.late_from = lf;
  return w;
}

// `at_cols(q)` (optional) is called after every step with the number q of leading columns whose block columns are
// final (panel factored and the rows below it solved): from then on L[:, 0:q] and the diagonal-block inverses of
// those columns no longer change.
template <class Hook>
static cudaError_t potrf_right_looking(const FactorCtx& c, int np, bool lookahead, Hook at_cols) {
  cudaError_t e;
  int NB = outer_nb();
  if (np % NB != 0) NB = BASE_N;
  const int nblk = np / NB;
  AuxStream* aux = nullptr;
  if (lookahead && nblk > 2) {
    if ((e = get_aux(&aux)) != cudaSuccess) return e;
  }
  const PanelWidths pw = panel_widths(np, NB, aux != nullptr);
  FactorCtx ca = c;                 // context of the panel (side) stream
  if (aux) ca.st = aux->st;
  auto at = [&](int i, int j) { return (long long)i * c.ld + j; };
  // first panel
  int c0 = 0, w0 = pw.at(0);
  if ((e = factor_rec(c, 0, w0)) != cudaSuccess) return e;
  if (w0 < np) {
    if ((e = gemm(true, true, c.A + at(w0, 0), c.ld, c.sb, c.M + at(0, 0), c.ld, c.sb, c.L + at(w0, 0), c.ld, c.sb,
                  np - w0, w0, w0, 1.0, 0.0, 0, K_LE_COL, c.batch, c.st)) != cudaSuccess) return e;
  }
  while (c0 + w0 < np) {
    // panel [c0, c1) is factored and solved; the next one is [c1, c2)
    const int c1 = c0 + w0, w1 = std::min(pw.at(c1), np - c1), c2 = c1 + w1;
    const int rows1 = np - c1;                      // rows below panel [c0, c1)
    // (a) rank-w0 update of block column [c1, c2)
    bool chain_released = false;
    g_trace_label = "a:colupd";
    if (!lookahead) {
      // batched problems (no panel chain to protect): the diagonal block of the column is symmetric - lower
      // tiles only, and only the lower fragments of its diagonal tiles - then the rectangle below it
      if ((e = gemm(true, true, c.L + at(c1, c0), c.ld, c.sb, c.L + at(c1, c0), c.ld, c.sb, c.A + at(c1, c1), c.ld,
                    c.sb, w1, w1, w0, -1.0, 1.0, 1, K_FULL, c.batch, c.st)) != cudaSuccess) return e;
      if (rows1 > w1 &&
          (e = gemm(true, true, c.L + at(c2, c0), c.ld, c.sb, c.L + at(c1, c0), c.ld, c.sb, c.A + at(c2, c1), c.ld,
                    c.sb, rows1 - w1, w1, w0, -1.0, 1.0, 0, K_FULL, c.batch, c.st)) != cudaSuccess) return e;
    } else {
      // look-ahead: the diagonal block of the column first - it is all the panel chain needs - then the rows below it,
      // which only the TRSM of the panel needs.  FFGP_SPLIT_COLUPD=1 enables it; measured at N = 8192: 20.9-21.6 ms/eval
      // against 20.6 with the single launch (the chain is not what bounds a step, profiles/r01_c2_schedule_experiments.txt)
      const int rows_first = (aux && split_colupd() && rows1 > w1) ? w1 : rows1;
      if ((e = gemm(true, true, c.L + at(c1, c0), c.ld, c.sb, c.L + at(c1, c0), c.ld, c.sb, c.A + at(c1, c1),
                    c.ld, c.sb, rows_first, w1, w0, -1.0, 1.0, 0, K_FULL, c.batch, c.st)) != cudaSuccess) return e;
      if (rows_first < rows1) {
        if ((e = cudaEventRecord(aux->ev_main, c.st)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(aux->st, aux->ev_main, 0)) != cudaSuccess) return e;
        chain_released = true;
        if ((e = gemm(true, true, c.L + at(c2, c0), c.ld, c.sb, c.L + at(c1, c0), c.ld, c.sb, c.A + at(c2, c1),
                      c.ld, c.sb, rows1 - w1, w1, w0, -1.0, 1.0, 0, K_FULL, c.batch, c.st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(aux->ev_rest, c.st)) != cudaSuccess) return e;
      }
    }
    cudaStream_t ps = c.st;
    if (aux) {
      if (!chain_released) {
        if ((e = cudaEventRecord(aux->ev_main, c.st)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(aux->st, aux->ev_main, 0)) != cudaSuccess) return e;
      }
      ps = aux->st;
    }
    // (b) panel [c1, c2): diagonal block (factor + inverse), then TRSM of the rows below it
    g_trace_label = "b:panel";
    if ((e = factor_rec(aux ? ca : c, c1, w1)) != cudaSuccess) return e;
    const int rows2 = np - c2;
    // Schedule 2 (FFGP_SCHED=2, experiment): the trailing update runs as a persistent grid that leaves FFGP_SYRK_RESERVE
    // SMs to the diagonal-block chain on the side stream, and the TRSM of the panel follows it on the MAIN stream (all
    // SMs).  The chain shrinks from ~200 to ~140 us per block but a step is bound by SYRK + TRSM + column update:
    // no gain at N = 8192 (profiles/r01_c2_schedule_experiments.txt), so the default stays schedule 1.
    const bool sched2 = aux && sched_mode() == 2;
    if (rows2 > 0) {
      if (!sched2) {
        g_trace_label = "b:trsm";
        if (chain_released && (e = cudaStreamWaitEvent(ps, aux->ev_rest, 0)) != cudaSuccess) return e;
        if ((e = gemm(true, true, c.A + at(c2, c1), c.ld, c.sb, c.M + at(c1, c1), c.ld, c.sb,
                      c.L + at(c2, c1), c.ld, c.sb, rows2, w1, w1, 1.0, 0.0, 0, K_LE_COL, c.batch, ps)) != cudaSuccess)
          return e;
      }
      // (c) rest of the trailing update with panel [c0, c1) (independent of the new panel): lower tiles of A[c2:, c2:]
      g_trace_label = "c:syrk";
      if (aux) g_reserve_sms = syrk_reserve();
      e = gemm(true, true, c.L + at(c2, c0), c.ld, c.sb, c.L + at(c2, c0), c.ld, c.sb, c.A + at(c2, c2), c.ld,
               c.sb, rows2, rows2, w0, -1.0, 1.0, 1, K_FULL, c.batch, c.st);
      g_reserve_sms = 0;
      if (e != cudaSuccess) return e;
    }
    if (aux) {
      if ((e = cudaEventRecord(aux->ev_aux, aux->st)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(c.st, aux->ev_aux, 0)) != cudaSuccess) return e;
    }
    if (sched2 && rows2 > 0) {
      g_trace_label = "b:trsm";
      if ((e = gemm(true, true, c.A + at(c2, c1), c.ld, c.sb, c.M + at(c1, c1), c.ld, c.sb,
                    c.L + at(c2, c1), c.ld, c.sb, rows2, w1, w1, 1.0, 0.0, 0, K_LE_COL, c.batch, c.st)) != cudaSuccess)
        return e;
    }
    if ((e = at_cols(c2)) != cudaSuccess) return e;
    c0 = c1; w0 = w1;
  }
  return cudaSuccess;
}
static cudaError_t potrf_right_looking(const FactorCtx& c, int np, bool lookahead) {
  return potrf_right_looking(c, np, lookahead, [](int) { return cudaSuccess; });
}

// M21 = -M22 L21 M11 for the node [off, off+n) split at h (general, sequential; used when np/NB is not a power of 2)
static cudaError_t trtri_rec(const FactorCtx& c, int off, int n, int NB) {
  if (n <= NB) return cudaSuccess;
  cudaError_t e;
  const int h = ((n / NB) + 1) / 2 * NB, m = n - h;
  if ((e = trtri_rec(c, off, h, NB)) != cudaSuccess) return e;
  if ((e = trtri_rec(c, off + h, m, NB)) != cudaSuccess) return e;
  const long long d0 = (long long)off * c.ld + off;
  const long long o21 = (long long)(off + h) * c.ld + off, o22 = (long long)(off + h) * c.ld + off + h;
  if ((e = gemm(true, false, c.M + o22, c.ld, c.sb, c.L + o21, c.ld, c.sb, c.A + o21, c.ld, c.sb, m, h, m, 1.0, 0.0, 0,
                K_LE_ROW, c.batch, c.st)) != cudaSuccess) return e;
  return gemm(true, false, c.A + o21, c.ld, c.sb, c.M + d0, c.ld, c.sb, c.M + o21, c.ld, c.sb, m, h, h, -1.0, 0.0, 0,
              K_GE_COL, c.batch, c.st);
}

static cudaError_t trtri_bottom_up(const FactorCtx& c, int np, int nb0 = 0) {
  int NB = nb0 > 0 ? nb0 : outer_nb();      // nb0: width of the diagonal blocks whose inverses are already complete
  if (np % NB != 0) NB = BASE_N;
  const int nblk = np / NB;
  if (nblk & (nblk - 1)) return trtri_rec(c, 0, np, NB);      // not a power of two: plain recursion
  cudaError_t e;
  for (int h = NB; h < np; h *= 2) {
    const int nodes = np / (2 * h);
    const long long node_stride = (long long)2 * h * (c.ld + 1);
    const long long o21 = (long long)h * c.ld, o22 = (long long)h * c.ld + h;
    // T = M22 L21 -> A21 (dead after the factorisation)
    if ((e = gemm(true, false, c.M + o22, c.ld, c.sb, c.L + o21, c.ld, c.sb, c.A + o21, c.ld, c.sb, h, h, h, 1.0, 0.0, 0,
                  K_LE_ROW, c.batch, c.st, nodes, node_stride, node_stride, node_stride)) != cudaSuccess) return e;
    // M21 = -T M11
    if ((e = gemm(true, false, c.A + o21, c.ld, c.sb, c.M, c.ld, c.sb, c.M + o21, c.ld, c.sb, h, h, h, -1.0, 0.0, 0,
                  K_GE_COL, c.batch, c.st, nodes, node_stride, node_stride, node_stride)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

static cudaError_t trtri_bottom_up_range(const FactorCtx& c, int off, int n, int nb0 = 0) {
  FactorCtx r = c;
  const long long d = (long long)off * c.ld + off;
  r.A += d; r.L += d; r.M += d;
  return trtri_bottom_up(r, n, nb0);
}

// Single large problem: factorisation AND triangular inverse with the first half of the inverse overlapped with the
// second half of the factorisation.  profiles/r01_launches_bench_v5.csv: from the matrix midpoint on, every
// right-looking step is bound by its serial panel chain (diag update -> two 128-blocks -> TRSM, ~230 us) while the
// trailing update needs less than that, so most SMs idle for ~3.6 ms at N = 8192; and the inverse
//   M21 = -M22 L21 M11  =  -M22 (L21 M11)
// has half of its FLOPs (inverse of the leading half, then W = L21 M11) depending only on the LEADING half of L,
// which is final at the midpoint.  Streams: panel chain (highest priority) > trailing updates > background W.
static int bg_cta_cap() {
  static int v = -1;
  // default 0 = no cap (measured at N=8192: bands of 98 CTAs 22.2 ms/eval vs 21.6 ms uncapped)
  if (v < 0) { const char* e = getenv("FFGP_BG_CAP"); v = e ? atoi(e) : 0; }
  return v;
}

// FFGP_BG_RESERVE = R > 0: the background products run as persistent grids on all SMs but R, which stay free for the
// panel chain and the trailing updates of the chain-bound steps (a 128 x 128 TMA GEMM CTA owns its SM until its tile retires)
static int bg_reserve() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_BG_RESERVE"); v = e ? atoi(e) : 0; }
  return v;
}

static cudaError_t factor_and_invert_overlapped(const FactorCtx& c, int np, int stop_after, bool* done) {
  *done = false;
  int NB = outer_nb();
  if (np % NB != 0) return cudaSuccess;
  const int nblk = np / NB;
  if (nblk < 8 || (nblk & (nblk - 1))) return cudaSuccess;          // needs a power-of-two block count >= 8
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("FFGP_OVERLAP"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
  if (!enabled) return cudaSuccess;
  cudaError_t e;
  AuxStream* aux = nullptr;
  if ((e = get_aux(&aux)) != cudaSuccess) return e;
  // fork: the factorisation runs on the library's middle-priority stream
  if ((e = cudaEventRecord(aux->ev_fork, c.st)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(aux->st_bulk, aux->ev_fork, 0)) != cudaSuccess) return e;
  FactorCtx cb = c; cb.st = aux->st_bulk;
  FactorCtx cg = c; cg.st = aux->st_bg;
  // Row-progressive inverse.  Hook points q = np/2, 3np/4, 7np/8, .. (FFGP_HOOKS of them, default 3): when the block
  // columns [0, q) of L are final, the rows [p, q) of M (p = the previous hook point) can be completed in the
  // background, and the part of the remaining rows' inverse that only needs L[:, 0:q) accumulated into X (kept in the
  // dead strictly-lower part of A):
  //   1. M[p:q, p:q]  bottom-up from the diagonal-block inverses
  //   2. M[p:q, 0:p]  = -M[p:q, p:q] X[p:q, 0:p]                                  (p > 0)
  //   3. X[q:np, 0:q] = L[q:np, 0:q] M[0:q, 0:q], in column chunks of KC: the chunk's own triangular block first
  //      (first touch), then its full-depth contribution to the columns on its left (beta = 1)
  // After the factorisation: M[p:np, p:np] bottom-up, M[p:np, 0:p] = -M[p:np, p:np] X[p:np, 0:p] for the last hook point p.
  // With the single hook at np/2 this is the round-1 schedule (X = W = L21 M11); the later hooks move ~1.7 ms of
  // products from the tail into the chain-bound last quarter of the factorisation, where most SMs idle.
  static int nhooks = -1;
  if (nhooks < 0) { const char* ev = getenv("FFGP_HOOKS"); nhooks = ev ? atoi(ev) : 3; if (nhooks < 1) nhooks = 1; }
  static int KCv = -1;                             // FFGP_BG_KC: K chunk of the background X products (256 / 512 / 1024)
  if (KCv < 0) { const char* ev = getenv("FFGP_BG_KC"); KCv = ev ? atoi(ev) : 512; if (KCv < 128 || KCv % 128) KCv = 512; }
  const int nb0 = panel_widths(np, NB, true).min_width();   // narrowest panel of the sweep: where the bottom-up inverse starts
  int p_done = 0;                                   // rows [0, p_done) of M are complete (or queued on the background stream)
  int hooks_left = nhooks;
  auto at_cols = [&](int q) -> cudaError_t {       // called when block columns [0, q) are final
    if (stop_after == 1 || hooks_left == 0) return cudaSuccess;
    const int rest = np - p_done;                  // next hook point: the midpoint of what is left
    if (q != p_done + rest / 2 || (rest / 2) % NB != 0 || rest / 2 < NB) return cudaSuccess;
    const int p = p_done;
    cudaError_t e2;
    if ((e2 = cudaEventRecord(aux->ev_half, cb.st)) != cudaSuccess) return e2;
    if ((e2 = cudaStreamWaitEvent(cg.st, aux->ev_half, 0)) != cudaSuccess) return e2;
    g_cta_cap = bg_cta_cap();
    g_reserve_sms = bg_reserve();
    g_trace_label = "bg:trtri";
    e2 = trtri_bottom_up_range(cg, p, q - p, nb0);
    if (p > 0 && e2 == cudaSuccess)
      e2 = gemm(true, false, c.M + (long long)p * c.ld + p, c.ld, c.sb, c.A + (long long)p * c.ld, c.ld, c.sb,
                c.M + (long long)p * c.ld, c.ld, c.sb, q - p, p, q - p, -1.0, 0.0, 0, K_LE_ROW, c.batch, cg.st);
    // one launch with K up to np/2 keeps every SM busy for ~100 us per tile and stalled the panel chain by 1.7 ms
    // (profiles/r01_timeline_c2_v1.txt, steps 19-20); with 512-deep tiles an SM frees up every ~7 us.
    g_trace_label = "bg:W";
    const long long oq = (long long)q * c.ld;      // row q of the arrays
    for (int s0 = p; s0 < q && e2 == cudaSuccess; s0 += KCv) {
      const int kc = std::min(KCv, q - s0);
      e2 = gemm(true, false, c.L + oq + s0, c.ld, c.sb, c.M + (long long)s0 * c.ld + s0, c.ld, c.sb, c.A + oq + s0, c.ld, c.sb,
                np - q, kc, kc, 1.0, 0.0, 0, K_GE_COL, c.batch, cg.st);
      if (s0 > 0 && e2 == cudaSuccess)
        e2 = gemm(true, false, c.L + oq + s0, c.ld, c.sb, c.M + (long long)s0 * c.ld, c.ld, c.sb, c.A + oq, c.ld, c.sb, np - q, s0,
                  kc, 1.0, 1.0, 0, K_FULL, c.batch, cg.st);
    }
    g_cta_cap = 0;
    g_reserve_sms = 0;
    p_done = q;
    --hooks_left;
    return e2;
  };
  if ((e = potrf_right_looking(cb, np, true, at_cols)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(aux->ev_join, cb.st)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(c.st, aux->ev_join, 0)) != cudaSuccess) return e;
  *done = true;
  if (stop_after == 1) return cudaSuccess;
  if (p_done == 0) { *done = false; return cudaErrorUnknown; }   // cannot happen: nblk >= 8 is a power of two
  // inverse of the trailing rows, then the one product that needs both parts
  g_trace_label = "post:trtri";
  const int p = p_done;
  if ((e = trtri_bottom_up_range(c, p, np - p, nb0)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(aux->ev_bg, cg.st)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(c.st, aux->ev_bg, 0)) != cudaSuccess) return e;
  return gemm(true, false, c.M + (long long)p * c.ld + p, c.ld, c.sb, c.A + (long long)p * c.ld, c.ld, c.sb,
              c.M + (long long)p * c.ld, c.ld, c.sb, np - p, p, np - p, -1.0, 0.0, 0, K_LE_ROW, c.batch, c.st);
}

// FFGP_DEBUG_STOP_AFTER=1|2|3 truncates an evaluation after potrf | trtri | the S = M^T M product (results are then
// meaningless): profiling aid for tools/phase_times.py, never set in production.
static int debug_stop_after() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FFGP_DEBUG_STOP_AFTER"); v = e ? atoi(e) : 0; }
  return v;
}

static cudaError_t ensure_attrs() {
  static PerDeviceOnce once;
  bool& g_attr_done = *once.slot();
  if (g_attr_done) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(potrf_trtri_base_kernel<BASE_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BASE_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(factor256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F2_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(potrf_trtri_base_kernel<BASE_N_BATCHED>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(symv_lower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 1024 * (int)sizeof(double));
  if (e != cudaSuccess) return e;
  for (auto kern : {grad_contract_kernel<false>, grad_contract_kernel<true>}) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grad_smem_doubles(GRAD_DMAX) * 8);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  g_attr_done = true;
  return cudaSuccess;
}

static size_t grad_smem_bytes(int d) {
  return grad_smem_doubles(d) * sizeof(double);
}

struct DenseArgs {
  const double *x, *y, *inv_ls, *amp, *diag_add, *sigma_add;
  int n, d, D, batch, params_batched, diag_batched, clamp;
};

// Assemble Sigma (lower) into ws.A and factor it: ws.L, ws.M, logdet partials.  Chunk-local batch nb.
static int assemble_and_factor(const DenseArgs& a, const DenseWs& w, int b0, int nb, int* info, cudaStream_t st) {
  KernelMatrixParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.x1 = a.x ? a.x + (long long)b0 * a.n * a.d : nullptr;
  kp.x2 = kp.x1;
  kp.sx1 = kp.sx2 = (long long)a.n * a.d;
  kp.w = a.inv_ls ? a.inv_ls + (a.params_batched ? (long long)b0 * a.d : 0) : nullptr;
  kp.sw = a.params_batched ? a.d : 0;
  kp.amp = a.amp ? a.amp + (a.params_batched ? b0 : 0) : nullptr;
  kp.samp = a.params_batched ? 1 : 0;
  kp.diag_add = a.diag_add ? a.diag_add + (a.diag_batched ? (long long)b0 * a.n : 0) : nullptr;
  kp.sdiag = a.diag_batched ? a.n : 0;
  kp.sigma_add = a.sigma_add ? a.sigma_add + (long long)b0 * a.n * a.n : nullptr;
  kp.ssig = (long long)a.n * a.n;
  kp.K = w.A; kp.n1 = a.n; kp.n2 = a.n; kp.d = a.d; kp.np1 = w.np; kp.np2 = w.np; kp.ldk = w.np;
  kp.sK = (long long)w.np * w.np;
  kp.symmetric = 1; kp.lower_only = 1; kp.clamp = a.clamp;
  trace_mark("start", 0, 0, st);
  // the plain kernel case (no covariance term, d <= 16) goes through the lower-super-tile kernel; FFGP_KM128=0: general kernel
  static int km128 = -1;
  if (km128 < 0) { const char* e = getenv("FFGP_KM128"); km128 = (e && atoi(e) == 0) ? 0 : 1; }
  if (km128 && kp.x1 && kp.amp && kp.w && !kp.sigma_add && a.d >= 1 && a.d <= 16 && nb <= 65535) {
    const int tt = w.np / 128;
    kernel_matrix_sym128_kernel<<<dim3(tt * (tt + 1) / 2, nb), 256, 0, st>>>(kp);
  } else {
    kernel_matrix_kernel<<<dim3(w.np / 64, w.np / 64, nb), 256, 0, st>>>(kp);
  }
  FFGP_LAUNCHED();
  trace_mark("kernel_matrix", 0, 0, st);
  FactorCtx c{w.A, w.L, w.M, w.np, (long long)w.np * w.np, nb, w.logdet_part, w.nblk, info + b0, st};
  c.base_n = (nb >= 8 && batched_base64()) ? BASE_N_BATCHED : BASE_N;
  // one log-det slot per 64 rows; with 128-blocks only every other slot is written
  FFGP_CUDA(cudaMemsetAsync(w.logdet_part, 0, sizeof(double) * (size_t)nb * w.nblk, st));
  // look-ahead needs spare SMs: with a large batch every launch already fills the machine
  if (nb < 8) {
    bool done = false;
    FFGP_CUDA(factor_and_invert_overlapped(c, w.np, debug_stop_after(), &done));
    if (done) return 0;
  }
  FFGP_CUDA(potrf_right_looking(c, w.np, /*lookahead=*/nb < 8));
  if (debug_stop_after() == 1) return 0;              // tools/phase_times.py: time the phases separately
  FFGP_CUDA(trtri_bottom_up(c, w.np, nb < 8 ? panel_widths(w.np, outer_nb(), true).min_width() : 0));
  return 0;
}

// Gamma, rowsq, alpha for the chunk (alpha stays padded in the workspace)
static int solve_rhs(const DenseArgs& a, const DenseWs& w, int b0, int nb, cudaStream_t st) {
  const long long sM = (long long)w.np * w.np;
  const double* y = a.y + (long long)b0 * a.n * a.D;
  if (!w.gemm_rhs) {
    const long long sG = (long long)w.np * a.D;
    trmv_lower_kernel<8><<<dim3(w.np / 8, nb), 256, 0, st>>>(w.M, w.np, sM, y, a.n, a.D, (long long)a.n * a.D, w.Gm, sG,
                                                             w.rowsq, w.np);
    FFGP_LAUNCHED();
    trace_mark("trmv_lower", 0, 0, st);
    colsum_weighted_kernel<8><<<dim3(w.np / 32, nb), 256, 0, st>>>(w.M, w.np, sM, w.np, w.Gm, a.D, sG, a.D, w.alpha, a.D,
                                                                   sG, w.np, 1, 0);
    FFGP_LAUNCHED();
    trace_mark("colsum_weighted", 0, 0, st);
  } else {
    const long long sG = (long long)w.np * w.Dp;
    dim3 blk(32, 8), grd((w.Dp + 31) / 32, (w.np + 7) / 8, nb);
    pad_copy_kernel<<<grd, blk, 0, st>>>(y, a.n, a.D, a.D, (long long)a.n * a.D, w.Ypad, w.np, w.Dp, w.Dp, sG);
    FFGP_LAUNCHED();
    trace_mark("pad_copy", 0, 0, st);
    FFGP_CUDA(gemm(true, false, w.M, w.np, sM, w.Ypad, w.Dp, sG, w.Gm, w.Dp, sG, w.np, w.Dp, w.np, 1.0, 0.0, 0, K_LE_ROW,
                   nb, st));
    rowsq_kernel<<<dim3(w.np / 8, nb), 256, 0, st>>>(w.Gm, w.Dp, sG, w.Dp, w.rowsq, w.np);
    FFGP_LAUNCHED();
    trace_mark("rowsq", 0, 0, st);
    FFGP_CUDA(gemm(false, false, w.M, w.np, sM, w.Gm, w.Dp, sG, w.alpha, w.Dp, sG, w.np, w.Dp, w.np, 1.0, 0.0, 0,
                   K_GE_ROW, nb, st));
  }
  return 0;
}

static int copy_alpha_out(const DenseArgs& a, const DenseWs& w, int b0, int nb, double* out_alpha, cudaStream_t st) {
  if (!out_alpha) return 0;
  const int Dw = w.gemm_rhs ? w.Dp : a.D;
  dim3 blk(32, 8), grd((a.D + 31) / 32, (a.n + 7) / 8, nb);
  pad_copy_kernel<<<grd, blk, 0, st>>>(w.alpha, w.np, Dw, Dw, (long long)w.np * Dw,
                                       out_alpha + (long long)b0 * a.n * a.D, a.n, a.D, a.D, (long long)a.n * a.D);
  FFGP_LAUNCHED();
  trace_mark("pad_copy", 0, 0, st);
  return 0;
}

}  // namespace ffgp

using namespace ffgp;

extern "C" {

int ffgp_version(void) { return FFGP_VERSION; }

// Debug only (FFGP_TRACE=1): synchronise and print "<ms since first mark> <stream> <label> <a> <b>" per recorded launch.
int ffgp_trace_dump(void) {
  if (!trace_on() || g_ntrace == 0) return 0;
  cudaDeviceSynchronize();
  for (int i = 0; i < g_ntrace; i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_trace[0].ev, g_trace[i].ev);
    printf("TRACE %9.3f %p %-12s %d %d\n", ms, g_trace[i].stream, g_trace[i].what, g_trace[i].a, g_trace[i].b);
  }
  for (int i = 0; i < g_ntrace; i++) cudaEventDestroy(g_trace[i].ev);
  const int n = g_ntrace;
  g_ntrace = 0;
  return n;
}
#ifdef FFGP_TG_TRACE
int ffgp_debug_tg_cta(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_tg_cta, sizeof(long long) * 4096 * 4); }
int ffgp_debug_tg_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_tg_trace, sizeof(long long) * 2 * 32 * 6); }
#endif
unsigned long long ffgp_launch_count(void) { return g_launches; }
const char* ffgp_last_error_string(void) { return g_err; }

int ffgp_gemm_f64(int a_kmajor, int b_kmajor, const double* A, int lda, long long strideA, const double* B, int ldb,
                  long long strideB, double* C, int ldc, long long strideC, int M, int N, int K, double alpha, double beta,
                  int lower_only, int kmode, int batch, void* stream) {
  if (!A || !B || !C) return fail(-1, "ffgp_gemm_f64: null pointer");
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0 || M % 64 || N % 64 || K % 16) return fail(-2, "ffgp_gemm_f64: M,N %% 64 and K %% 16 required");
  if (lower_only && M != N) return fail(-2, "ffgp_gemm_f64: lower_only needs M == N");
  if (kmode < 0 || kmode > 4) return fail(-2, "ffgp_gemm_f64: bad kmode");
  if ((lda | ldb | ldc) & 1) return fail(-2, "ffgp_gemm_f64: leading dimensions must be even (16-byte rows)");
  FFGP_CUDA(gemm(a_kmajor != 0, b_kmajor != 0, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, M, N, K, alpha, beta,
                 lower_only, kmode, batch, (cudaStream_t)stream));
  return 0;
}

size_t ffgp_dense_workspace_bytes(int n, int d, int D, int ns, int batch) {
  if (n <= 0 || D <= 0 || batch <= 0 || d < 0) return 0;
  return layout_ws(n, d, D, ns, batch, nullptr).bytes;
}

int ffgp_kernel_matrix_f64(const double* x1, const double* x2, const double* inv_ls, const double* amp, int n1, int n2,
                           int d, int batch, int params_batched, int clamp, double* K, void* stream) {
  if (!x1 || !x2 || !inv_ls || !amp || !K) return fail(-1, "ffgp_kernel_matrix_f64: null pointer");
  if (n1 <= 0 || n2 <= 0 || d <= 0 || batch <= 0) return fail(-2, "ffgp_kernel_matrix_f64: bad size");
  cudaStream_t st = (cudaStream_t)stream;
  KernelMatrixParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.x1 = x1; kp.x2 = x2; kp.sx1 = (long long)n1 * d; kp.sx2 = (long long)n2 * d;
  kp.w = inv_ls; kp.sw = params_batched ? d : 0; kp.amp = amp; kp.samp = params_batched ? 1 : 0;
  kp.K = K; kp.n1 = n1; kp.n2 = n2; kp.d = d;
  kp.np1 = round_up(n1, 64); kp.np2 = round_up(n2, 64); kp.ldk = n2; kp.sK = (long long)n1 * n2;
  kp.symmetric = 0; kp.lower_only = 0; kp.clamp = clamp; kp.bounded = 1;
  kernel_matrix_kernel<<<dim3(kp.np2 / 64, kp.np1 / 64, batch), 256, 0, st>>>(kp);
  FFGP_LAUNCHED();
  trace_mark("kernel_matrix", 0, 0, st);
  return 0;
}

int ffgp_dense_fit_f64(const double* x, const double* y, const double* xs, const double* inv_ls, const double* amp,
                       const double* diag_add, const double* sigma_add, const double* Ks, const double* Kss,
                       const double* cov_offset, int n, int d, int D, int ns, int batch, int params_batched, int clamp,
                       int want_nll, int want_grad, int full_cov, int reuse_factor, void* workspace, size_t workspace_bytes,
                       double* out_nll, double* out_logdet, double* out_alpha, double* g_inv_ls, double* g_amp,
                       double* g_diag, double* g_sigma, double* out_mean, double* out_cov, int* info, void* stream) {
  const bool want_pred = out_mean != nullptr;
  if (!y || !workspace || !info) return fail(-1, "ffgp_dense_fit_f64: null pointer");
  if (want_nll && !out_nll) return fail(-1, "ffgp_dense_fit_f64: out_nll missing");
  if (amp && (!x || !inv_ls)) return fail(-1, "ffgp_dense_fit_f64: kernel term needs x and inv_ls");
  if (!amp && !sigma_add) return fail(-1, "ffgp_dense_fit_f64: neither a kernel nor a covariance was given");
  if (n <= 0 || D <= 0 || batch <= 0 || d < 0 || ns < 0) return fail(-2, "ffgp_dense_fit_f64: bad size");
  if (amp && d > GRAD_DMAX && want_grad) return fail(-2, "ffgp_dense_fit_f64: d > 128 not supported with want_grad");
  if (want_grad && amp && (!g_inv_ls || !g_amp)) return fail(-1, "ffgp_dense_fit_f64: gradient outputs missing");
  if (want_pred) {
    if (ns <= 0) return fail(-2, "ffgp_dense_fit_f64: ns must be positive when predictions are requested");
    if (amp && !xs) return fail(-1, "ffgp_dense_fit_f64: kernel term needs xs for predictions");
    if (!amp && !Ks) return fail(-1, "ffgp_dense_fit_f64: covariance mode needs Ks");
    if (!amp && out_cov && !Kss) return fail(-1, "ffgp_dense_fit_f64: covariance mode needs Kss for out_cov");
    if (!amp && out_cov && !full_cov) return fail(-2, "ffgp_dense_fit_f64: covariance mode returns the full covariance");
  }
  cudaStream_t st = (cudaStream_t)stream;
  DenseWs w = layout_ws(n, d, D, want_pred ? ns : 0, batch, (char*)workspace);
  if (workspace_bytes < w.bytes) return fail(-3, "ffgp_dense_fit_f64: workspace too small");
  if (reuse_factor && batch > w.chunk) return fail(-4, "ffgp_dense_fit_f64: reuse_factor needs batch <= one chunk");
  if (reuse_factor && (want_nll || want_grad)) return fail(-4, "ffgp_dense_fit_f64: reuse_factor is for prediction only");
  FFGP_CUDA(ensure_attrs());
  DenseArgs a{x, y, inv_ls, amp, diag_add, sigma_add, n, d, D, batch, params_batched, /*diag_batched=*/params_batched, clamp};
  if (!reuse_factor) FFGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * batch, st));
  const long long sM = (long long)w.np * w.np;
  struct PersistGuard { ~PersistGuard() { g_persist = 0; } } persist_guard;
  // One chunk (problems [b0, b0 + nb) in the workspace slice `w`) on stream `st`.
  auto run_chunk = [&](int b0, int nb, const DenseWs& w, cudaStream_t st) -> int {
    int rc;
    // Batched gradient evaluations need S = Sigma^-1 anyway: alpha = S y in ONE pass over S (symv_lower_kernel) replaces
    // Gamma = M y and alpha = M^T Gamma (two passes over M) and y^T alpha gives the quadratic form.  FFGP_SYMV=0: old path.
    static int symv_on = -1;
    if (symv_on < 0) { const char* e = getenv("FFGP_SYMV"); symv_on = (e && atoi(e) == 0) ? 0 : 1; }
    const bool use_symv = symv_on && want_grad && !reuse_factor && !w.gemm_rhs && nb >= 8 && w.np <= 1024 &&
                          debug_stop_after() == 0;
    if (!reuse_factor) {
      if ((rc = assemble_and_factor(a, w, b0, nb, info, st)) != 0) return rc;
      if (debug_stop_after() == 1 || debug_stop_after() == 2) return 0;
      if (!use_symv && (rc = solve_rhs(a, w, b0, nb, st)) != 0) return rc;
    }
    if (use_symv) {
      FFGP_CUDA(gemm(false, false, w.M, w.np, sM, w.M, w.np, sM, w.A, w.np, sM, w.np, w.np, w.np, 1.0, 0.0, 1, K_GE_ROW, nb, st));
      const long long sG = (long long)w.np * D;
      symv_lower_kernel<<<dim3(nb, D), 256, (size_t)9 * w.np * sizeof(double), st>>>(
          w.A, w.np, sM, n, w.np, y + (long long)b0 * n * D, D, (long long)n * D, w.alpha, sG, w.rowsq);
      FFGP_LAUNCHED();
      trace_mark("symv_lower", 0, 0, st);
      if (D > 1) {
        rowdot_kernel<<<dim3((w.np + 255) / 256, nb), 256, 0, st>>>(y + (long long)b0 * n * D, (long long)n * D, w.alpha, sG, n,
                                                                     w.np, D, w.rowsq);
        FFGP_LAUNCHED();
        trace_mark("rowdot", 0, 0, st);
      }
    }
    if (want_nll) {
      nll_reduce_kernel<<<nb, 256, 0, st>>>(w.rowsq, w.np, w.logdet_part, w.nblk, D, out_nll + b0,
                                            out_logdet ? out_logdet + b0 : nullptr);
      FFGP_LAUNCHED();
      trace_mark("nll_reduce", 0, 0, st);
    }
    if (!reuse_factor && (rc = copy_alpha_out(a, w, b0, nb, out_alpha, st)) != 0) return rc;
    if (want_grad) {
      // S = M^T M (lower) into the dead A buffer
      if (!use_symv)
        FFGP_CUDA(gemm(false, false, w.M, w.np, sM, w.M, w.np, sM, w.A, w.np, sM, w.np, w.np, w.np, 1.0, 0.0, 1, K_GE_ROW, nb, st));
      if (debug_stop_after() == 3) return 0;
      int src_is_G = 0;
      if (w.gemm_rhs) {   // G = 0.5 (D S - alpha alpha^T) through the GEMM epilogue
        const long long sG = (long long)w.np * w.Dp;
        FFGP_CUDA(gemm(true, true, w.alpha, w.Dp, sG, w.alpha, w.Dp, sG, w.A, w.np, sM, w.np, w.np, w.Dp, -0.5, 0.5 * D, 1,
                       K_FULL, nb, st));
        src_is_G = 1;
      }
      GradParams gp;
      memset(&gp, 0, sizeof(gp));
      gp.src = w.A; gp.ld = w.np; gp.ssrc = sM;
      gp.x = x ? x + (long long)b0 * n * d : nullptr; gp.n = n; gp.d = amp ? d : 0; gp.sx = (long long)n * d;
      gp.w = inv_ls ? inv_ls + (params_batched ? (long long)b0 * d : 0) : nullptr; gp.sw = params_batched ? d : 0;
      gp.amp = amp ? amp + (params_batched ? b0 : 0) : nullptr; gp.samp = params_batched ? 1 : 0;
      gp.alpha = w.alpha; gp.D = D; gp.salpha = (long long)w.np * D;
      gp.src_is_G = src_is_G;
      gp.partial = w.partial; gp.npart = w.ngtile;
      gp.g_diag = g_diag ? g_diag + (long long)b0 * n : nullptr; gp.sgd = n;
      gp.G_out = g_sigma ? g_sigma + (long long)b0 * n * n : nullptr; gp.sGo = (long long)n * n;
      gp.have_k = amp ? 1 : 0;
      if (gp.G_out) grad_contract_kernel<true><<<dim3(w.ngtile, nb), 256, grad_smem_bytes(gp.d), st>>>(gp);
      else grad_contract_kernel<false><<<dim3(w.ngtile, nb), 256, grad_smem_bytes(gp.d), st>>>(gp);
      FFGP_LAUNCHED();
      trace_mark("grad_contract", 0, 0, st);
      if (amp) {
        grad_finish_kernel<<<dim3(d + 1, nb), 256, 0, st>>>(w.partial, w.ngtile, d, gp.w, gp.sw, gp.amp, gp.samp,
                                                            g_inv_ls + (long long)b0 * d, g_amp + b0);
        FFGP_LAUNCHED();
        trace_mark("grad_finish", 0, 0, st);
      }
    }
    if (!want_pred) return 0;
    // ---------------- posterior at xs, from the factor that is still resident (M, alpha) ----------------
    const long long sKx = (long long)w.np * w.nsp, sKxx = (long long)w.nsp * w.nsp;
    KernelMatrixParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.n1 = n; kp.n2 = ns; kp.d = d; kp.np1 = w.np; kp.np2 = w.nsp; kp.ldk = w.nsp; kp.sK = sKx; kp.K = w.Kx;
    kp.clamp = clamp;
    if (amp) {
      kp.x1 = x + (long long)b0 * n * d; kp.sx1 = (long long)n * d;
      kp.x2 = xs + (long long)b0 * ns * d; kp.sx2 = (long long)ns * d;
      kp.w = inv_ls + (params_batched ? (long long)b0 * d : 0); kp.sw = params_batched ? d : 0;
      kp.amp = amp + (params_batched ? b0 : 0); kp.samp = params_batched ? 1 : 0;
    } else {
      kp.sigma_add = Ks + (long long)b0 * n * ns; kp.ssig = (long long)n * ns;
    }
    kernel_matrix_kernel<<<dim3(w.nsp / 64, w.np / 64, nb), 256, 0, st>>>(kp);
    FFGP_LAUNCHED();
    trace_mark("kernel_matrix", 0, 0, st);
    // mean = Kx^T alpha
    if (!w.gemm_rhs) {
      colsum_weighted_kernel<8><<<dim3(w.nsp / 32, nb), 256, 0, st>>>(w.Kx, w.nsp, sKx, w.np, w.alpha, D,
                                                                      (long long)w.np * D, D,
                                                                      out_mean + (long long)b0 * ns * D, D,
                                                                      (long long)ns * D, ns, 0, 0);
      FFGP_LAUNCHED();
      trace_mark("colsum_weighted", 0, 0, st);
    } else {
      const long long sG = (long long)w.np * w.Dp, sMp = (long long)w.nsp * w.Dp;
      FFGP_CUDA(gemm(false, false, w.Kx, w.nsp, sKx, w.alpha, w.Dp, sG, w.meanp, w.Dp, sMp, w.nsp, w.Dp, w.np, 1.0, 0.0, 0,
                     K_FULL, nb, st));
      dim3 blk(32, 8), grd((D + 31) / 32, (ns + 7) / 8, nb);
      pad_copy_kernel<<<grd, blk, 0, st>>>(w.meanp, w.nsp, w.Dp, w.Dp, sMp, out_mean + (long long)b0 * ns * D, ns, D, D,
                                           (long long)ns * D);
      FFGP_LAUNCHED();
      trace_mark("pad_copy", 0, 0, st);
    }
    if (!out_cov) return 0;
    // V = M Kx
    FFGP_CUDA(gemm(true, false, w.M, w.np, sM, w.Kx, w.nsp, sKx, w.V, w.nsp, sKx, w.np, w.nsp, w.np, 1.0, 0.0, 0, K_LE_ROW,
                   nb, st));
    if (full_cov) {
      KernelMatrixParams kq;
      memset(&kq, 0, sizeof(kq));
      kq.n1 = ns; kq.n2 = ns; kq.d = d; kq.np1 = w.nsp; kq.np2 = w.nsp; kq.ldk = w.nsp; kq.sK = sKxx; kq.K = w.Kxx;
      kq.clamp = clamp;
      if (amp) {
        kq.x1 = kq.x2 = xs + (long long)b0 * ns * d; kq.sx1 = kq.sx2 = (long long)ns * d;
        kq.w = kp.w; kq.sw = kp.sw; kq.amp = kp.amp; kq.samp = kp.samp;
      } else {
        kq.sigma_add = Kss + (long long)b0 * ns * ns; kq.ssig = (long long)ns * ns;
      }
      kq.offset = cov_offset ? cov_offset + (params_batched ? b0 : 0) : nullptr; kq.soff = params_batched ? 1 : 0;
      kernel_matrix_kernel<<<dim3(w.nsp / 64, w.nsp / 64, nb), 256, 0, st>>>(kq);
      FFGP_LAUNCHED();
      trace_mark("kernel_matrix", 0, 0, st);
      // cov = Kxx - V^T V
      FFGP_CUDA(gemm(false, false, w.V, w.nsp, sKx, w.V, w.nsp, sKx, w.Kxx, w.nsp, sKxx, w.nsp, w.nsp, w.np, -1.0, 1.0, 0,
                     K_FULL, nb, st));
      dim3 blk(32, 8), grd((ns + 31) / 32, (ns + 7) / 8, nb);
      unpad_copy_kernel<<<grd, blk, 0, st>>>(w.Kxx, w.nsp, sKxx, out_cov + (long long)b0 * ns * ns, ns, ns,
                                             (long long)ns * ns, 0);
      FFGP_LAUNCHED();
      trace_mark("unpad_copy", 0, 0, st);
    } else {
      colsum_weighted_kernel<1><<<dim3(w.nsp / 32, nb), 256, 0, st>>>(w.V, w.nsp, sKx, w.np, nullptr, 0, 0, 1, w.colsq, 1,
                                                                      w.nsp, w.nsp, 0, 1);
      FFGP_LAUNCHED();
      trace_mark("colsum_weighted", 0, 0, st);
      var_diag_kernel<<<dim3((ns + 127) / 128, nb), 128, 0, st>>>(w.colsq, amp + (params_batched ? b0 : 0),
                                                                  params_batched ? 1 : 0,
                                                                  cov_offset ? cov_offset + (params_batched ? b0 : 0) : nullptr,
                                                                  params_batched ? 1 : 0, out_cov + (long long)b0 * ns, ns, w.nsp);
      FFGP_LAUNCHED();
      trace_mark("var_diag", 0, 0, st);
    }
    return 0;
  };
  // Two streams for one chunk (FFGP_SPLIT):
  //   2 (default)  tail split: only the problems of the last, partial wave of the one-CTA-per-problem launches (nb mod SM
  //                count) go to the library's second stream, so that their fused 256-block CTAs overlap the first part's next
  //                product instead of leaving most SMs idle for a whole wave.  512 problems per GPU (BASELINE config 5 on 8
  //                GPUs) = 3.46 waves of 148: 5.485 -> 5.29 ms per sweep (-3.6 %, same results bit for bit); 4096 problems
  //                (tail 100): 38.10 vs 38.12 ms, neutral; 1024 / 2048 problems (tails 136 / 124: last wave > 3/4 full) are
  //                not split - every split point measured there is neutral or slower (profiles/r02_c5_experiments.txt).
  //   1            the two HALVES of a chunk on two streams, so that the ramp-up / drain of the ~30 dependent launches of one
  //                half is filled by the other half's CTAs: 41.56 ms per sweep against 41.07 on one stream - every GEMM CTA
  //                owns a whole SM, so the second chain only ever gets the SMs the first one has drained.  Not adopted.
  //   0            one stream.
  static int split_on = -1;
  if (split_on < 0) { const char* e = getenv("FFGP_SPLIT"); split_on = e ? atoi(e) : 2; if (split_on < 0 || split_on > 2) split_on = 2; }
  for (int b0 = 0; b0 < batch; b0 += w.chunk) {
    const int nb = std::min(w.chunk, batch - b0);
    g_persist = nb >= 8;      // same rule as the look-ahead switch: a large batch fills the machine by itself
    int rc;
    AuxStream* aux = nullptr;
    const int tail = nb % num_sms();
    if (((split_on == 1 && nb >= 4 * num_sms()) || (split_on == 2 && nb >= 2 * num_sms() && tail >= 8 && tail <= num_sms() * 3 / 4)) &&
        !reuse_factor && debug_stop_after() == 0 && get_aux(&aux) == cudaSuccess) {
      const int sms = num_sms();
      // mode 2: whole waves first; below four waves the two parts are made comparable (512 problems: 296 + 216 measured
      // 5.290 ms, 444 + 68: 5.366, 148 + 364: 5.323, one stream: 5.485), above that only the partial wave moves
      int na = split_on == 2 ? (nb < 4 * sms ? ((nb / sms) + 1) / 2 * sms : nb - tail)
                             : std::min(nb - sms, (nb / 2 + sms - 1) / sms * sms);
      static int na_override = -1;                  // FFGP_SPLIT_NA: size of the first part (experiments)
      if (na_override < 0) { const char* e = getenv("FFGP_SPLIT_NA"); na_override = e ? atoi(e) : 0; }
      if (na_override > 0 && na_override < nb) na = na_override;
      FFGP_CUDA(cudaEventRecord(aux->ev_fork, st));
      FFGP_CUDA(cudaStreamWaitEvent(aux->st_bulk, aux->ev_fork, 0));
      if ((rc = run_chunk(b0, na, w, st)) != 0) return rc;
      if ((rc = run_chunk(b0 + na, nb - na, ws_offset(w, na, d, D), aux->st_bulk)) != 0) return rc;
      FFGP_CUDA(cudaEventRecord(aux->ev_join, aux->st_bulk));
      FFGP_CUDA(cudaStreamWaitEvent(st, aux->ev_join, 0));
    } else {
      if ((rc = run_chunk(b0, nb, w, st)) != 0) return rc;
    }
  }
  return 0;
}

int ffgp_dense_nll_f64(const double* x, const double* y, const double* inv_ls, const double* amp, const double* diag_add,
                       const double* sigma_add, int n, int d, int D, int batch, int params_batched, int clamp,
                       int want_grad, void* workspace, size_t workspace_bytes, double* out_nll, double* out_logdet,
                       double* out_alpha, double* g_inv_ls, double* g_amp, double* g_diag, double* g_sigma, int* info,
                       void* stream) {
  return ffgp_dense_fit_f64(x, y, nullptr, inv_ls, amp, diag_add, sigma_add, nullptr, nullptr, nullptr, n, d, D, 0, batch,
                            params_batched, clamp, 1, want_grad, 0, 0, workspace, workspace_bytes, out_nll, out_logdet,
                            out_alpha, g_inv_ls, g_amp, g_diag, g_sigma, nullptr, nullptr, info, stream);
}

int ffgp_dense_predict_f64(const double* x, const double* y, const double* xs, const double* inv_ls, const double* amp,
                           const double* diag_add, const double* sigma_add, const double* Ks, const double* Kss,
                           const double* cov_offset, int n, int d, int D, int ns, int batch, int params_batched,
                           int clamp, int full_cov, int reuse_factor, void* workspace, size_t workspace_bytes,
                           double* out_mean, double* out_cov, int* info, void* stream) {
  if (!out_mean) return fail(-1, "ffgp_dense_predict_f64: null pointer");
  return ffgp_dense_fit_f64(x, y, xs, inv_ls, amp, diag_add, sigma_add, Ks, Kss, cov_offset, n, d, D, ns, batch,
                            params_batched, clamp, 0, 0, full_cov, reuse_factor, workspace, workspace_bytes, nullptr,
                            nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, out_mean, out_cov, info, stream);
}

int ffgp_potrf_trtri_f64(const double* A, int n, int batch, void* workspace, size_t workspace_bytes, double* L,
                         double* Linv, double* logdet, int* info, void* stream) {
  if (!A || !workspace || !info) return fail(-1, "ffgp_potrf_trtri_f64: null pointer");
  if (n <= 0 || batch <= 0) return fail(-2, "ffgp_potrf_trtri_f64: bad size");
  cudaStream_t st = (cudaStream_t)stream;
  DenseWs w = layout_ws(n, 0, 1, 0, batch, (char*)workspace);
  if (workspace_bytes < w.bytes) return fail(-3, "ffgp_potrf_trtri_f64: workspace too small");
  FFGP_CUDA(ensure_attrs());
  DenseArgs a{nullptr, nullptr, nullptr, nullptr, nullptr, A, n, 0, 1, batch, 0, 0, 0};
  FFGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * batch, st));
  for (int b0 = 0; b0 < batch; b0 += w.chunk) {
    const int nb = std::min(w.chunk, batch - b0);
    int rc;
    if ((rc = assemble_and_factor(a, w, b0, nb, info, st)) != 0) return rc;
    dim3 blk(32, 8), grd((n + 31) / 32, (n + 7) / 8, nb);
    if (L) {
      unpad_copy_kernel<<<grd, blk, 0, st>>>(w.L, w.np, (long long)w.np * w.np, L + (long long)b0 * n * n, n, n, (long long)n * n, 1);
      FFGP_LAUNCHED();
    }
    if (Linv) {
      unpad_copy_kernel<<<grd, blk, 0, st>>>(w.M, w.np, (long long)w.np * w.np, Linv + (long long)b0 * n * n, n, n, (long long)n * n, 1);
      FFGP_LAUNCHED();
    }
    if (logdet) {
      // rowsq is unused here: zero it so the reducer returns logdet only
      FFGP_CUDA(cudaMemsetAsync(w.rowsq, 0, sizeof(double) * (size_t)nb * w.np, st));
      nll_reduce_kernel<<<nb, 256, 0, st>>>(w.rowsq, w.np, w.logdet_part, w.nblk, 1, nullptr, logdet + b0);
      FFGP_LAUNCHED();
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Input-point gradients (acquisition optimisers: d posterior / d x*)
// ---------------------------------------------------------------------------------------------
static size_t xgrad_part_bytes(int R, int C, int d, int batch) {
  return (size_t)batch * xgrad_splits(R, C, batch) * R * d * sizeof(double);
}

static int launch_xgrad(XGradParams p, int batch, double* out, int accumulate, cudaStream_t st) {
  p.splits = xgrad_splits(p.R, p.C, batch);
  xgrad_kernel<<<dim3((p.R + XG_WARPS - 1) / XG_WARPS, p.splits, batch), XG_WARPS * 32, 0, st>>>(p);
  FFGP_LAUNCHED();
  const long long rd = (long long)p.R * p.d;
  xgrad_finish_kernel<<<dim3((unsigned)((rd + 255) / 256), batch), 256, 0, st>>>(p.part, p.splits, rd, out, accumulate);
  FFGP_LAUNCHED();
  return 0;
}

size_t ffgp_kernel_matrix_bwd_x_scratch_bytes(int n1, int n2, int d, int batch) {
  if (n1 <= 0 || n2 <= 0 || d <= 0 || batch <= 0) return 0;
  return std::max(xgrad_part_bytes(n1, n2, d, batch), xgrad_part_bytes(n2, n1, d, batch)) + 256;
}

int ffgp_kernel_matrix_bwd_x_f64(const double* x1, const double* x2, const double* inv_ls, const double* amp,
                                 const double* gK, int n1, int n2, int d, int batch, int params_batched, double* g_x1,
                                 double* g_x2, void* scratch, size_t scratch_bytes, void* stream) {
  if (!x1 || !x2 || !inv_ls || !amp || !gK || !scratch) return fail(-1, "ffgp_kernel_matrix_bwd_x_f64: null pointer");
  if (n1 <= 0 || n2 <= 0 || d <= 0 || d > XG_DMAX || batch <= 0) return fail(-2, "ffgp_kernel_matrix_bwd_x_f64: bad size (d <= 64)");
  if (scratch_bytes < ffgp_kernel_matrix_bwd_x_scratch_bytes(n1, n2, d, batch)) return fail(-3, "ffgp_kernel_matrix_bwd_x_f64: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  XGradParams p;
  memset(&p, 0, sizeof(p));
  p.w = inv_ls; p.sw = params_batched ? d : 0; p.amp = amp; p.samp = params_batched ? 1 : 0;
  p.gK = gK; p.sgK = (long long)n1 * n2; p.gk_scale = 1.0; p.d = d; p.part = (double*)scratch;
  int rc;
  if (g_x1) {
    p.xr = x1; p.xc = x2; p.sxr = (long long)n1 * d; p.sxc = (long long)n2 * d; p.R = n1; p.C = n2; p.ldr = n2; p.ldc = 1;
    if ((rc = launch_xgrad(p, batch, g_x1, 0, st)) != 0) return rc;
  }
  if (g_x2) {
    p.xr = x2; p.xc = x1; p.sxr = (long long)n2 * d; p.sxc = (long long)n1 * d; p.R = n2; p.C = n1; p.ldr = 1; p.ldc = n2;
    if ((rc = launch_xgrad(p, batch, g_x2, 0, st)) != 0) return rc;
  }
  return 0;
}

size_t ffgp_dense_predict_bwd_scratch_bytes(int n, int d, int ns, int batch) {
  if (n <= 0 || ns <= 0 || d <= 0 || batch <= 0) return 0;
  return std::max(xgrad_part_bytes(ns, n, d, batch), xgrad_part_bytes(ns, ns, d, batch)) + 256;
}

int ffgp_dense_predict_bwd_f64(const double* x, const double* xs, const double* inv_ls, const double* amp,
                               const double* g_mean, const double* g_cov, int n, int d, int D, int ns, int batch,
                               int params_batched, int full_cov, void* workspace, size_t workspace_bytes, double* g_xs,
                               void* scratch, size_t scratch_bytes, void* stream) {
  if (!x || !xs || !inv_ls || !amp || !workspace || !g_xs || !scratch) return fail(-1, "ffgp_dense_predict_bwd_f64: null pointer");
  if (!g_mean && !g_cov) return fail(-1, "ffgp_dense_predict_bwd_f64: neither g_mean nor g_cov given");
  if (n <= 0 || D <= 0 || batch <= 0 || d <= 0 || d > XG_DMAX || ns <= 0) return fail(-2, "ffgp_dense_predict_bwd_f64: bad size (d <= 64)");
  if (scratch_bytes < ffgp_dense_predict_bwd_scratch_bytes(n, d, ns, batch)) return fail(-3, "ffgp_dense_predict_bwd_f64: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  DenseWs w = layout_ws(n, d, D, ns, batch, (char*)workspace);
  if (workspace_bytes < w.bytes) return fail(-3, "ffgp_dense_predict_bwd_f64: workspace too small");
  if (batch > w.chunk) return fail(-4, "ffgp_dense_predict_bwd_f64: batch must fit one workspace chunk (the factor must be resident)");
  const long long sM = (long long)w.np * w.np, sKx = (long long)w.np * w.nsp, sKxx = (long long)w.nsp * w.nsp;
  const int Dw = w.gemm_rhs ? w.Dp : D;
  XGradParams p;
  memset(&p, 0, sizeof(p));
  p.w = inv_ls; p.sw = params_batched ? d : 0; p.amp = amp; p.samp = params_batched ? 1 : 0; p.d = d;
  p.part = (double*)scratch;
  p.xr = xs; p.sxr = (long long)ns * d; p.R = ns;
  p.xc = x; p.sxc = (long long)n * d; p.C = n;
  if (g_mean) {                                    // d mean / d K* = alpha g_mean^T
    p.lr_c = w.alpha; p.ld_lrc = Dw; p.s_lrc = (long long)w.np * Dw; p.lr_r = g_mean; p.lrD = D; p.s_lrr = (long long)ns * D;
  }
  int rc;
  if (g_cov) {
    // R = M^T V = Sigma^-1 K*  (into the dead K* buffer; the forward recomputes K* on every call)
    FFGP_CUDA(gemm(false, false, w.M, w.np, sM, w.V, w.nsp, sKx, w.Kx, w.nsp, sKx, w.np, w.nsp, w.np, 1.0, 0.0, 0, K_GE_ROW,
                   batch, st));
    const double* T = w.Kx;
    if (full_cov) {                                // d cov / d K* = -R (G + G^T)
      dim3 blk(32, 8), grd((w.nsp + 31) / 32, (w.nsp + 7) / 8, batch);
      sym_pad_kernel<<<grd, blk, 0, st>>>(g_cov, ns, (long long)ns * ns, w.Kxx, w.nsp, sKxx);
      FFGP_LAUNCHED();
      FFGP_CUDA(gemm(true, false, w.Kx, w.nsp, sKx, w.Kxx, w.nsp, sKxx, w.V, w.nsp, sKx, w.np, w.nsp, w.nsp, 1.0, 0.0, 0,
                     K_FULL, batch, st));
      T = w.V;
    } else {                                       // d var_j / d K*[:, j] = -2 R[:, j] g_var[j]
      p.colscale = g_cov;
    }
    p.gK = T; p.ldr = 1; p.ldc = w.nsp; p.sgK = sKx; p.gk_scale = full_cov ? -1.0 : -2.0;
  }
  if ((rc = launch_xgrad(p, batch, g_xs, 0, st)) != 0) return rc;
  if (g_cov && full_cov) {                         // + d K(x*, x*) / d x*  with the symmetrised weights
    XGradParams q;
    memset(&q, 0, sizeof(q));
    q.w = p.w; q.sw = p.sw; q.amp = p.amp; q.samp = p.samp; q.d = d; q.part = (double*)scratch;
    q.xr = xs; q.xc = xs; q.sxr = q.sxc = (long long)ns * d; q.R = ns; q.C = ns;
    q.gK = w.Kxx; q.ldr = w.nsp; q.ldc = 1; q.sgK = sKxx; q.gk_scale = 1.0;
    if ((rc = launch_xgrad(q, batch, g_xs, 1, st)) != 0) return rc;
  }
  return 0;
}

static AcqConsts acq_consts(int kind, double f_best, double beta, double xi, int round_f32) {
  AcqConsts c;
  c.kind = kind; c.f_best = f_best; c.beta = beta; c.xi = xi;
  c.std_min = 1e-9; c.two_pi = 2.0 * 3.1415926;          /* DMF_acq.py:7,98,121; acq.py:178,226 */
  c.round_f32 = round_f32;
  return c;
}

int ffgp_acquisition_f64(const double* mean, const double* var, int m, int kind, double f_best, double beta, double xi,
                         int round_f32, double* score, double* d_mean, double* d_var, void* stream) {
  if (!mean || !var || !score) return fail(-1, "ffgp_acquisition_f64: null pointer");
  if (m <= 0 || kind < 0 || kind > 4)
    return fail(-2, "ffgp_acquisition_f64: bad size or kind (0 UCB_MF, 1 EI, 2 PI_MF, 3 UCB on the std, 4 PI as cdf)");
  AcqParams p;
  p.mean = mean; p.var = var; p.m = m;
  p.c = acq_consts(kind, f_best, beta, xi, round_f32);
  p.score = score; p.d_mean = d_mean; p.d_var = d_var;
  acq_kernel<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
  FFGP_LAUNCHED();
  return 0;
}


int ffgp_adam_step_f64(void* const* table, const int* sizes, int ntensors, double lr, double beta1, double beta2,
                       double eps, int maximize, const double* loss, double* loss_hist, int hist_cap, void* stream) {
  if (!table || !sizes) return fail(-1, "ffgp_adam_step_f64: null pointer");
  if (ntensors <= 0) return fail(-2, "ffgp_adam_step_f64: ntensors must be positive");
  if (!(lr >= 0.0) || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0))
    return fail(-2, "ffgp_adam_step_f64: bad hyper-parameter");
  adam_step_kernel<<<ntensors, 256, 0, (cudaStream_t)stream>>>(table, sizes, lr, beta1, beta2, eps, maximize, loss, loss_hist,
                                                             loss_hist ? hist_cap : 0);
  FFGP_LAUNCHED();
  return 0;
}

int ffgp_batched_pack_acq_f64(const double* nll_core, const double* g_inv_ls, const double* g_amp, const double* g_diag,
                              const double* mean, const double* var, const int* info, const double* length_scales,
                              const double* signal_variance, const double* log_beta, int batch, int n, int d, int D, int ns,
                              int want_grad, int with_info, double nll_const, double eps, int acq_kind, double f_best,
                              double beta, double xi, int round_f32, double* out, int ld_out, void* stream) {
  if (!nll_core || !out) return fail(-1, "ffgp_batched_pack_f64: null pointer");
  if (want_grad && (!g_inv_ls || !g_amp || !g_diag || !length_scales || !signal_variance || !log_beta))
    return fail(-1, "ffgp_batched_pack_f64: gradient inputs missing");
  if (ns > 0 && (!mean || !var)) return fail(-1, "ffgp_batched_pack_f64: prediction inputs missing");
  if (with_info && !info) return fail(-1, "ffgp_batched_pack_f64: info missing");
  if (batch <= 0 || n <= 0 || d < 0 || D <= 0 || ns < 0) return fail(-2, "ffgp_batched_pack_f64: bad size");
  if (acq_kind > 4) return fail(-2, "ffgp_batched_pack_acq_f64: bad acquisition kind");
  if (acq_kind >= 0 && (ns <= 0 || D != 1)) return fail(-2, "ffgp_batched_pack_acq_f64: scores need test points and D = 1");
  const int need = 1 + (want_grad ? d + 2 : 0) + (ns > 0 ? ns * D + ns : 0) + (acq_kind >= 0 ? ns : 0) + (with_info ? 1 : 0);
  if (ld_out < need) return fail(-2, "ffgp_batched_pack_f64: ld_out too small");
  PackParams p;
  p.nll_core = nll_core; p.g_il = g_inv_ls; p.g_amp = g_amp; p.g_diag = g_diag; p.mean = mean; p.var = var; p.info = info;
  p.ls = length_scales; p.sv = signal_variance; p.lb = log_beta;
  p.B = batch; p.n = n; p.d = d; p.D = D; p.ns = ns; p.want_grad = want_grad; p.with_info = with_info;
  p.nll_const = nll_const; p.eps = eps; p.out = out; p.ld = ld_out;
  p.acq = acq_consts(acq_kind < 0 ? -1 : acq_kind, f_best, beta, xi, round_f32);
  pack_results_kernel<<<batch, 128, 0, (cudaStream_t)stream>>>(p);
  FFGP_LAUNCHED();
  return 0;
}

int ffgp_batched_pack_f64(const double* nll_core, const double* g_inv_ls, const double* g_amp, const double* g_diag,
                          const double* mean, const double* var, const int* info, const double* length_scales,
                          const double* signal_variance, const double* log_beta, int batch, int n, int d, int D, int ns,
                          int want_grad, int with_info, double nll_const, double eps, double* out, int ld_out, void* stream) {
  return ffgp_batched_pack_acq_f64(nll_core, g_inv_ls, g_amp, g_diag, mean, var, info, length_scales, signal_variance, log_beta,
                                   batch, n, d, D, ns, want_grad, with_info, nll_const, eps, -1, 0.0, 0.0, 0.0, 0, out, ld_out,
                                   stream);
}

int ffgp_row_match_f64(const double* a, const double* b, int na, int nb, int d, int* match, void* stream) {
  if (!a || !b || !match) return fail(-1, "ffgp_row_match_f64: null pointer");
  if (na <= 0 || nb <= 0 || d <= 0) return fail(-2, "ffgp_row_match_f64: bad size");
  const size_t smem = (size_t)MATCH_ROWS * (2 * d + 1) * sizeof(double);
  if (smem > 200 * 1024) return fail(-2, "ffgp_row_match_f64: d too large (rows of up to 99 values)");
  static PerDeviceOnce once;
  bool& attr = *once.slot();
  if (!attr) {
    FFGP_CUDA(cudaFuncSetAttribute(row_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  row_match_kernel<<<(na + MATCH_ROWS - 1) / MATCH_ROWS, MATCH_ROWS, smem, (cudaStream_t)stream>>>(a, b, na, nb, d, match);
  FFGP_LAUNCHED();
  return 0;
}

}  // extern "C"
