/* libffgp - B200 (sm_100a) Gaussian-process hot path behind FidelityFusion's operator API.
 *
 * C ABI (plain pointers and sizes, no torch types).  Every pointer is a DEVICE pointer unless
 * marked host.  The caller owns every buffer, including the opaque workspace whose size is
 * returned by the *_workspace_bytes queries; the library allocates nothing and keeps no pointer
 * after a call returns.  All work is enqueued on `stream` (a cudaStream_t passed as void*) and
 * there is no host synchronisation inside the library.  Matrices are row-major fp64.
 *
 * Return value: 0 = enqueued; <0 = bad argument / CUDA error (see ffgp_last_error_string()).
 * Numerical failure (leading minor not positive definite) is reported LAPACK-style through the
 * device array `info[batch]` (0 = ok, k>0 = minor k), which the host checks after the stream
 * has drained - the Python layer turns it into torch.linalg.LinAlgError, which is what
 * torch.linalg.cholesky raises in the reference.
 *
 * The reference has no FFI layer of its own (it is pure PyTorch); each entry point below
 * replaces the chain of torch calls cited next to it.  Paths are relative to the reference
 * root (IceLab-X/FidelityFusion).
 */
#ifndef FFGP_H
#define FFGP_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFGP_VERSION 100

int ffgp_version(void);
const char* ffgp_last_error_string(void);  /* host string, thread-local */
int ffgp_trace_dump(void);                  /* debug: with FFGP_TRACE=1, print the recorded launch timeline (host sync) */
int ffgp_debug_last_eigh_sweeps(void);      /* debug: Jacobi sweeps of the most recent ffgp_syevj_f64 solve with n <= 128 (host sync) */
unsigned long long ffgp_launch_count(void); /* kernels launched by the library so far (host counter) */

/* ---------------------------------------------------------------------------------------
 * The level-3 building block of the whole dense path, exported for unit tests and profiling:
 *   C = alpha * op(A) * op(B) + beta * C   on the FP64 tensor pipe (DMMA), row-major, `batch` strided problems.
 *   a_kmajor: A(i,p) = A[i*lda+p] else A[p*lda+i];   b_kmajor: B(p,j) = B[j*ldb+p] else B[p*ldb+j]
 *   lower_only: only tiles on/below the diagonal are produced (M == N);
 *   kmode: 0 full K; 1 p < i0+T (A lower-tri rows); 2 p < j0+T (B[j][p] lower-tri); 3 p >= j0 (B[p][j] lower-tri);
 *          4 p >= i0 (A[p][i] lower-tri) - the K range of a tile shrinks, no special kernel.
 * M, N must be multiples of 64 and K of 16 (the library pads its own problems to 128).
 * Replaces torch.triangular_solve / cholesky_solve / mm / LinalgCholeskyExBackward0's GEMMs.
 * --------------------------------------------------------------------------------------- */
int ffgp_gemm_f64(int a_kmajor, int b_kmajor, const double* A, int lda, long long strideA,
                  const double* B, int ldb, long long strideB, double* C, int ldc, long long strideC,
                  int M, int N, int K, double alpha, double beta, int lower_only, int kmode, int batch, void* stream);

/* ---------------------------------------------------------------------------------------
 * Stationary squared-exponential family, one parameterisation for the reference's three:
 *   K[i][j] = amp * exp(-0.5 * sum_k ((x1[i][k] - x2[j][k]) * inv_ls[k])^2)
 *   ARDKernel                 GaussianProcess/kernel.py:100-105   inv_ls = 1/(|p|+eps), amp = |s|, clamp=1 (cdist)
 *   SquaredExponentialKernel  GaussianProcess/kernel.py:271-272   inv_ls = exp(-l),     amp = exp(s)^2
 *   SE_kernel                 MFGP_ver2023May/kernel/SE_kernel.py:20-44  inv_ls = exp(-p) or 1/p, amp = exp(q) or q
 * x1 [batch][n1][d], x2 [batch][n2][d], K [batch][n1][n2].  inv_ls [d] and amp [1] are shared by
 * the batch when params_batched == 0, else [batch][d] / [batch].
 * --------------------------------------------------------------------------------------- */
int ffgp_kernel_matrix_f64(const double* x1, const double* x2, const double* inv_ls, const double* amp,
                           int n1, int n2, int d, int batch, int params_batched, int clamp,
                           double* K, void* stream);

/* d(sum(gK o K))/d(inv_ls, amp) for the kernel above (autograd of a stand-alone kernel call).
 * gK [batch][n1][n2]; g_inv_ls [batch][d]; g_amp [batch].  scratch: ffgp_kernel_matrix_bwd_scratch_bytes(). */
size_t ffgp_kernel_matrix_bwd_scratch_bytes(int n1, int n2, int d, int batch);
int ffgp_kernel_matrix_bwd_f64(const double* x1, const double* x2, const double* inv_ls, const double* amp,
                               const double* gK, int n1, int n2, int d, int batch, int params_batched,
                               double* g_inv_ls, double* g_amp, void* scratch, size_t scratch_bytes, void* stream);

/* d(sum(gK o K))/d(x1, x2): gradient w.r.t. the INPUT points of the kernel above (either output may be NULL).
 * Replaces autograd's CdistBackward0/PowBackward0/ExpBackward0 chain of ARDKernel.forward (kernel.py:100-105) and the
 * MmBackward chain of the two SE kernels when a caller differentiates K w.r.t. its inputs (acquisition optimisers,
 * DMF_acq.py:226-262).  g_x1 [batch][n1][d], g_x2 [batch][n2][d]. */
size_t ffgp_kernel_matrix_bwd_x_scratch_bytes(int n1, int n2, int d, int batch);
int ffgp_kernel_matrix_bwd_x_f64(const double* x1, const double* x2, const double* inv_ls, const double* amp,
                                 const double* gK, int n1, int n2, int d, int batch, int params_batched,
                                 double* g_x1, double* g_x2, void* scratch, size_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dense GP negative log marginal likelihood (+ analytic gradient), `batch` independent problems.
 *   Sigma = K(x,x) + diag(diag_add) + sigma_add ;  L = chol(Sigma) ;  Gamma = L^-1 y
 *   out_nll[b] = 0.5*||Gamma||_F^2 + D * sum_i log L_ii          (the 0.5*n*D*log(2*pi') constant is
 *                                                                  added by the caller: the reference
 *                                                                  uses 3.1415 or math.pi per path)
 * Replaces: cigp.negative_log_likelihood  GaussianProcess/cigp_v10.py:50-69
 *           CIGP.compute_loss             MFGP_ver2023May/base_gp/cigp.py:99-136
 *           Gaussian_log_likelihood       GaussianProcess/gp_computation_pack.py:34-91 (amp == NULL: Sigma = sigma_add)
 *           GP_basic.log_likelihood       GaussianProcess/gp_basic.py:94-153
 *           and, with want_grad, their autograd backward (LinalgCholeskyExBackward0, TriangularSolveBackward0,
 *           CdistBackward0 ...) through the closed form G = dNLL/dSigma = 0.5*(D*Sigma^-1 - alpha alpha^T).
 * Inputs : x [batch][n][d]; y [batch][n][D]; inv_ls/amp as above (amp == NULL => no kernel term);
 *          diag_add [n] or [batch][n] or NULL (noise + jitter + diag(y_var)); sigma_add [batch][n][n] or NULL.
 * Outputs: out_nll [batch]; out_logdet [batch] or NULL (= 2 sum log L_ii);
 *          out_alpha [batch][n][D] = Sigma^-1 y (= dNLL/dy ... with sign: dNLL/dy = alpha);
 *          with want_grad: g_inv_ls [batch][d], g_amp [batch], g_diag [batch][n] (= G_ii),
 *          g_sigma [batch][n][n] or NULL (full G, for callers that differentiate a given covariance).
 *          info [batch] int.
 * --------------------------------------------------------------------------------------- */
size_t ffgp_dense_workspace_bytes(int n, int d, int D, int ns, int batch);

int ffgp_dense_nll_f64(const double* x, const double* y, const double* inv_ls, const double* amp,
                       const double* diag_add, const double* sigma_add,
                       int n, int d, int D, int batch, int params_batched, int clamp, int want_grad,
                       void* workspace, size_t workspace_bytes,
                       double* out_nll, double* out_logdet, double* out_alpha,
                       double* g_inv_ls, double* g_amp, double* g_diag, double* g_sigma,
                       int* info, void* stream);

/* ---------------------------------------------------------------------------------------
 * Posterior mean / covariance at xs [batch][ns][d].
 *   V = L^-1 K(x,xs);  mean = K(x,xs)^T Sigma^-1 y;  cov = K(xs,xs) - V^T V + cov_offset
 * Replaces: cigp.forward            GaussianProcess/cigp_v10.py:24-48   (full_cov=1, cov_offset = e^-log_beta on EVERY entry)
 *           CIGP.forward            MFGP_ver2023May/base_gp/cigp.py:61-97 (full_cov=0: diagonal + 1/beta)
 *           conditional_Gaussian    GaussianProcess/gp_computation_pack.py:93-118 (amp == NULL: K_s, K_ss given)
 *           GP_basic.forward        GaussianProcess/gp_basic.py:40-92
 * reuse_factor != 0: the workspace still holds L^-1 and alpha of the same training problem from the
 * previous ffgp_dense_nll_f64 / ffgp_dense_predict_f64 call (factor caching);
 * batch must then fit in one workspace chunk.
 * amp == NULL: Ks [batch][n][ns] and Kss [batch][ns][ns] (or NULL when only the mean is wanted) are given.
 * out_mean [batch][ns][D]; out_cov [batch][ns][ns] (full_cov) or [batch][ns] (diagonal).
 * --------------------------------------------------------------------------------------- */
int ffgp_dense_predict_f64(const double* x, const double* y, const double* xs,
                           const double* inv_ls, const double* amp, const double* diag_add, const double* sigma_add,
                           const double* Ks, const double* Kss, const double* cov_offset,
                           int n, int d, int D, int ns, int batch, int params_batched, int clamp,
                           int full_cov, int reuse_factor,
                           void* workspace, size_t workspace_bytes,
                           double* out_mean, double* out_cov, int* info, void* stream);

/* ---------------------------------------------------------------------------------------
 * Gradient of the posterior w.r.t. the test points:  g_xs = d( <g_mean, mean> + <g_cov, cov> ) / d xs.
 * Replaces the autograd backward of cigp.forward (cigp_v10.py:24-48) that the acquisition optimisers run per
 * candidate step (DMF_acq.py:247-254, v1/MF_EI.py:22-31): closed form
 *   d/dK* = alpha g_mean^T - Sigma^-1 K* (G + G^T)      (diagonal variance: - 2 Sigma^-1 K* diag(g_var))
 * followed by the input-point gradient of K* (and of K(xs,xs) for the full covariance).
 * MUST be called with the workspace of the ffgp_dense_predict_f64 call it differentiates (same n, d, D, ns, batch,
 * out_cov requested when g_cov is given): it reads L^-1, alpha and V from it and overwrites K*, V, Kxx.
 * g_mean [batch][ns][D] or NULL; g_cov [batch][ns][ns] (full_cov) / [batch][ns] or NULL; g_xs [batch][ns][d].
 * Kernel mode only (amp != NULL); the hyper-parameters are treated as constants.
 * --------------------------------------------------------------------------------------- */
size_t ffgp_dense_predict_bwd_scratch_bytes(int n, int d, int ns, int batch);
int ffgp_dense_predict_bwd_f64(const double* x, const double* xs, const double* inv_ls, const double* amp,
                               const double* g_mean, const double* g_cov, int n, int d, int D, int ns, int batch,
                               int params_batched, int full_cov, void* workspace, size_t workspace_bytes,
                               double* g_xs, void* scratch, size_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Acquisition scores of m candidates from their posterior mean / variance, with the partial derivatives
 * d score / d mean and d score / d var (either may be NULL) in the same pass - chained with
 * ffgp_dense_predict_bwd_f64 this is the whole candidate-optimisation step on the device.
 * Replaces DiscreteAcquisitionFunction.UCB_MF / EI_MF / PI_MF (MF_BayesianOptimization/Discrete/DMF_acq.py:47-128),
 * whose EI goes through scipy.stats.norm on the host (DMF_acq.py:104).
 *   kind 0: mean + beta * var            kind 1: t Phi(z) + s phi(z), t = mean - f_best - xi, s = max(sqrt(var), 1e-9), z = t / s
 *   kind 2: -z^2/2 - log sqrt(2 * 3.1415926)
 * and the single-fidelity classes of Bayesian_optimization/acq.py (UCB :135-149, PI :211-231; its EI :176-181 is kind 1):
 *   kind 3: mean + beta * sqrt(var)      kind 4: Phi(z) (float32-rounded when round_f32; a constant: both partials are 0)
 * round_f32 != 0 (EI): Phi and phi are rounded to float32 as the reference's torch.tensor(norm.cdf(..), dtype=float32).
 * --------------------------------------------------------------------------------------- */
int ffgp_acquisition_f64(const double* mean, const double* var, int m, int kind, double f_best, double beta, double xi,
                         int round_f32, double* score, double* d_mean, double* d_var, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused fit: NLL (+ gradient) AND the posterior at xs from ONE factorisation per problem - what a BO acquisition
 * sweep needs per candidate (train step, then predict: v1/CFKG.py:124-129 re-trains and predicts per candidate).
 * Same arguments as ffgp_dense_nll_f64 + ffgp_dense_predict_f64; out_mean == NULL skips the prediction,
 * want_nll == 0 skips the NLL.  ffgp_dense_nll_f64 / ffgp_dense_predict_f64 are thin wrappers over it.
 * --------------------------------------------------------------------------------------- */
int ffgp_dense_fit_f64(const double* x, const double* y, const double* xs, const double* inv_ls, const double* amp,
                       const double* diag_add, const double* sigma_add, const double* Ks, const double* Kss,
                       const double* cov_offset, int n, int d, int D, int ns, int batch, int params_batched, int clamp,
                       int want_nll, int want_grad, int full_cov, int reuse_factor,
                       void* workspace, size_t workspace_bytes,
                       double* out_nll, double* out_logdet, double* out_alpha,
                       double* g_inv_ls, double* g_amp, double* g_diag, double* g_sigma,
                       double* out_mean, double* out_cov, int* info, void* stream);

/* ---------------------------------------------------------------------------------------
 * Packed result rows of a batch of cigp + ARDKernel problems (the buffer ONE all-gather ships):
 *   out[b] = [ nll | d/d length_scales [d] | d/d signal_variance | d/d log_beta | mean [ns*D] | var [ns] | info ]
 * from the outputs of ffgp_dense_fit_f64 (nll_core, g_inv_ls, g_amp, g_diag, mean, diagonal var, info).  The chain rule to
 * the reference's raw parameters (inv_ls = 1/(|length_scales| + eps), amp = |signal_variance|, diag = e^-log_beta + jitter:
 * GaussianProcess/kernel.py:100-105, cigp_v10.py:57-58) and nll = nll_core + nll_const are applied in the same launch.
 * The gradient block is present iff want_grad, the prediction block iff ns > 0, the status column iff with_info.
 * --------------------------------------------------------------------------------------- */
int ffgp_batched_pack_f64(const double* nll_core, const double* g_inv_ls, const double* g_amp, const double* g_diag,
                          const double* mean, const double* var, const int* info, const double* length_scales,
                          const double* signal_variance, const double* log_beta, int batch, int n, int d, int D, int ns,
                          int want_grad, int with_info, double nll_const, double eps, double* out, int ld_out, void* stream);

/* The same row with the acquisition score of every test point appended after the variances (D = 1, ns > 0):
 *   [.. | mean[ns] | var[ns] | score[ns] | info],  score = ffgp_acquisition_f64's (acq_kind, f_best, beta, xi, round_f32).
 * The erfc / exp epilogue of the batched posterior: the candidates' scores of an acquisition sweep
 * (MF_BayesianOptimization/Discrete/DMF_acq.py:82-104, Bayesian_optimization/acq.py:118-256) leave the device-side sweep in
 * the same row - and through the same all-gather - as the predictions; the reference round-trips mean and variance through
 * scipy.stats.norm on the host per candidate set (DMF_acq.py:104).  acq_kind < 0: exactly ffgp_batched_pack_f64. */
int ffgp_batched_pack_acq_f64(const double* nll_core, const double* g_inv_ls, const double* g_amp, const double* g_diag,
                              const double* mean, const double* var, const int* info, const double* length_scales,
                              const double* signal_variance, const double* log_beta, int batch, int n, int d, int D, int ns,
                              int want_grad, int with_info, double nll_const, double eps, int acq_kind, double f_best,
                              double beta, double xi, int round_f32, double* out, int ld_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Subset / overlap matching of two fidelities' inputs: match[i] = smallest j with b[j][:] == a[i][:]
 * under IEEE equality (NaN matches nothing, -0.0 == 0.0), else -1.  a [na][d], b [nb][d], match int[na].
 * Replaces the [na][nb][d] broadcast compare of FidelityFusion_Models/MF_data.py:199-202 / 235-238
 * (get_overlap_input_data / get_unique_input_data) and the unique()-based MFGP_ver2023May/utils/subset_tools.py:58-90.
 * --------------------------------------------------------------------------------------- */
int ffgp_row_match_f64(const double* a, const double* b, int na, int nb, int d, int* match, void* stream);

/* ---------------------------------------------------------------------------------------
 * Device-side optimiser step.  torch.optim.Adam semantics (no weight decay, no amsgrad; operation
 * order of torch/optim/adam.py::_single_tensor_adam) for `ntensors` parameter tensors in ONE launch; replaces the
 * optimizer.step() of the reference's training loops (FidelityFusion_Models/CIGAR.py:96-106, AR_autoRegression.py:
 * 120-140, GaussianProcess/cigp_v10.py:160-175, MFGP_ver2023May/mfgp_demo.py:40-48).
 *   table  device array of 5*ntensors pointers: {param, grad, exp_avg, exp_avg_sq, step} per tensor (fp64, contiguous;
 *          step is a device double[1] per tensor, zero before the first update and advanced by the kernel, so a
 *          captured CUDA graph of one epoch replays with no host value changing between iterations)
 *   sizes  device int[ntensors] element counts
 *   loss / loss_hist (optional): loss_hist[step-1] = loss[0] for the first hist_cap steps (training curve without a
 *          device-to-host read per iteration)
 * --------------------------------------------------------------------------------------- */
int ffgp_adam_step_f64(void* const* table, const int* sizes, int ntensors, double lr, double beta1, double beta2,
                       double eps, int maximize, const double* loss, double* loss_hist, int hist_cap, void* stream);

/* ---------------------------------------------------------------------------------------
 * Cholesky factor and triangular inverse of `batch` SPD matrices (row-major, lower).
 * Replaces torch.linalg.cholesky + L.inverse() (cigp.py:129-131, gp_computation_pack.py:108-109).
 * A [batch][n][n] (only the lower triangle is read); L, Linv [batch][n][n] (either may be NULL).
 * --------------------------------------------------------------------------------------- */
int ffgp_potrf_trtri_f64(const double* A, int n, int batch, void* workspace, size_t workspace_bytes,
                         double* L, double* Linv, double* logdet, int* info, void* stream);

/* ---------------------------------------------------------------------------------------
 * n-mode product (tensorly.tenalg.mode_dot; call sites hogp.py:132,181,183,187,217,236,
 * multiscale_coupling/matrix.py:73,81, gp_computation_pack.py:157):
 *   out[.., j, ..] = sum_i mat[j][i] * t[.., i, ..]  along `mode` of a contiguous tensor viewed as
 *   [outer][I][inner];  mat is [J][I] (transpose_mat == 0) or [I][J] (transpose_mat != 0).
 * --------------------------------------------------------------------------------------- */
int ffgp_mode_dot_f64(const double* t, const double* mat, double* out,
                      long long outer, int I, long long inner, int J, int transpose_mat, void* stream);

/* ---------------------------------------------------------------------------------------
 * Symmetric eigendecomposition of `batch` small matrices (cyclic Jacobi in shared memory),
 * ascending eigenvalues like torch.linalg.eigh(K, UPLO='U') (hogp.py:18-22).  n <= 512.
 * A [batch][n][n] is read from the UPPER triangle; w [batch][n]; V [batch][n][n] (columns = vectors).
 * --------------------------------------------------------------------------------------- */
size_t ffgp_syevj_workspace_bytes(int n, int batch);
int ffgp_syevj_f64(const double* A, int n, int batch, double* w, double* V,
                   void* workspace, size_t workspace_bytes, int* info, void* stream);

/* G[a][b] = sum_{outer,inner} X[o][a][in] * Y[o][b][in]  (X viewed [outer][Ja][inner], Y [outer][Jb][inner]).
 * This is d(mode_dot)/d(mat) (autograd of the couplings trained through the residual,
 * multiscale_coupling/matrix.py:64, gp_computation_pack.py:152) and the weighted mode Gram matrices of the
 * Kronecker-GP gradient.  Split over column slices, summed in a fixed order (deterministic). */
size_t ffgp_mode_gram_scratch_bytes(long long outer, long long inner, int Ja, int Jb);
int ffgp_mode_gram_f64(const double* X, const double* Y, double* G, long long outer, long long inner, int Ja, int Jb,
                       void* scratch, size_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Kronecker / Tucker GP core stage (HOGP.compute_loss hogp.py:171-198, HOGP_simple.log_likelihood
 * hogp_simple.py:104-126): with per-mode eigenvalues lambda_k (concatenated, `sizes_host[nmodes]` on the HOST),
 *   A = kron(lambda_0..lambda_M) + noise_inv[0] + add_scalar,   T1 = Y x_k U_k^T (computed by ffgp_mode_dot_f64)
 * one pass over the tensor produces
 *   out_core = T1 / A   (then g = out_core x_k U_k),  out_A (optional),
 *   out_sums[4] = { sum log A, sum T1^2 / A, sum 1/A, sum (T1/A)^2 }   (fixed-order reduction).
 * ffgp_kron_scale_f64: out = in o prod_{m != skip_mode} lambda_m[i_m]  (/ A when divide_by_A) - the weights
 * of the analytic gradient w.r.t. each mode's kernel matrix; in == NULL means in = 1.
 * --------------------------------------------------------------------------------------- */
size_t ffgp_kron_core_scratch_bytes(long long total);
int ffgp_kron_core_f64(const double* T1, const double* lambdas, const int* sizes_host, int nmodes,
                       const double* noise_inv, double add_scalar,
                       double* out_core, double* out_A, double* out_sums, void* scratch, size_t scratch_bytes,
                       void* stream);
int ffgp_kron_scale_f64(const double* in, const double* lambdas, const int* sizes_host, int nmodes, int skip_mode,
                        int divide_by_A, const double* noise_inv, double add_scalar, double* out, void* stream);
/* c_k[j] = sum over the other modes' indices of prod_{m != mode} lambda_m / A (the diagonal weights of the closed-form
 * gradient of the Kronecker objective w.r.t. mode `mode`'s kernel matrix, replacing autograd through torch.linalg.eigh in
 * hogp.py:18-22,171-198): a reduction over the eigenvalues only, no pass over the data tensor.  out[sizes[mode]]. */
size_t ffgp_kron_ck_scratch_bytes(int n_k);
int ffgp_kron_ck_f64(const double* lambdas, const int* sizes_host, int nmodes, int mode, const double* noise_inv,
                     double add_scalar, double* out, void* scratch, size_t scratch_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FFGP_H */
