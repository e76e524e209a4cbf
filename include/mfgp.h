/* mfgp.h -- C-ABI of libmfgp.so: the B200 (sm_100a) implementation of the Kennedy-O'Hagan
 * linear multi-fidelity GP objective, its gradient and its predictive equations.
 *
 * Drop-in boundary.  The reference (qezlou/multi_fidelity_gpflow, pure Python) reaches
 * this arithmetic through GPflow's operator API; each entry point below names the
 * reference interface it replaces (file:line in /root/reference).  INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - All matrices are dense, row-major, float64.  `X` arrays are [N, d+1] with the
 *    fidelity indicator (exactly 0.0 or 1.0) in the last column (linear.py:67-70); rows
 *    with any other fidelity value have all-zero covariance (linear.py:82,99-102).
 *  - theta (per kernel) = [rho, ls_L[0..d-1], var_L, ls_delta[0..d-1], var_delta],
 *    length 2d+3, CONSTRAINED values (the softplus chain rule lives in the host shim).
 *  - Every pointer argument may be a HOST pointer or a DEVICE pointer of the handle's
 *    device; the library detects which (cudaPointerGetAttributes) and stages host
 *    buffers through the handle's stream.  Caller owns all buffers.
 *  - Return value: 0 ok; >0 = LAPACK-style info (1-based index of the first non-positive
 *    Cholesky pivot; for batched calls the first failing problem's info, per-problem
 *    values in `info[]`); <0 = bad argument (-1), CUDA error (-2), not supported (-3).
 *    mfgp_last_error() returns the message.
 *  - A handle is not thread-safe: one handle per GPU per thread.  Calls are ordered on
 *    the handle's stream and, unless async mode is on, synchronous on return.
 */
#ifndef MFGP_H
#define MFGP_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfgp_handle mfgp_handle;

/* ---- lifetime / control ------------------------------------------------------------ */
int mfgp_version(void);
int mfgp_create(int device, mfgp_handle** out);
int mfgp_destroy(mfgp_handle* h);
/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL is the legacy default
 * stream).  mfgp_reset_stream() returns to the handle's own non-blocking stream. */
int mfgp_set_stream(mfgp_handle* h, void* cuda_stream);
int mfgp_reset_stream(mfgp_handle* h);
/* async != 0: calls whose outputs are all device pointers return without synchronising;
 * non-PD status is then collected by mfgp_sync(). */
int mfgp_set_async(mfgp_handle* h, int async);
int mfgp_sync(mfgp_handle* h, int* info_out);
const char* mfgp_last_error(mfgp_handle* h);
int mfgp_sm_count(mfgp_handle* h);

/* ---- K1: covariance assembly ---------------------------------------------------------
 * replaces LinearMultiFidelityKernel.K (mfgpflow/linear.py:55-104) and K_diag (:106-136)
 * X2 == NULL -> X2 = X (linear.py:59-60): the symmetric fast path computes the lower
 * triangle of tiles once and mirrors it.  K is [N, N2] with leading dimension ldk. */
int mfgp_cov(mfgp_handle* h, const double* X, int N, const double* X2, int N2, int d,
             const double* theta, double* K, long ldk);
int mfgp_cov_diag(mfgp_handle* h, const double* X, int N, int d, const double* theta, double* out);

/* K5: out[q] = scale * sum_ij G_ij dK_ij/dtheta_q (q < 2d+3) with K = K(X, X; theta) and dK recomputed on the fly (never
 * stored); out[2d+3] = scale * sum_i G_ii (derivative of the noise term).  G [N, ldg] is symmetric and only its LOWER triangle
 * is read (off-diagonal entries count twice).  This is the contraction of tape.gradient through K (linear.py:207) exposed for
 * callers that build G = alpha alpha^T - K^-1 themselves (the distributed exact-GP gradient, dist_chol.py). */
int mfgp_cov_grad(mfgp_handle* h, const double* X, int N, int d, const double* theta, const double* G, long ldg, double scale,
                  double* out /* [2d+4] */);

/* ---- K2..K5: exact multi-fidelity GPR -------------------------------------------------
 * replaces gpflow GPR.log_marginal_likelihood as called at linear.py:206,227 (model built
 * at linear.py:148-156: zero mean, Gaussian noise `noise`, ONE kernel/Cholesky shared by
 * all P columns of Y [N, P]).  *nlml = -log_marginal_likelihood. */
int mfgp_gpr_nlml(mfgp_handle* h, const double* X, const double* Y, int N, int d, int P,
                  const double* theta, double noise, double* nlml);
/* + analytic gradient of nlml w.r.t. [theta (2d+3), noise] -- replaces
 * tape.gradient(loss, trainable_variables) at linear.py:207 (constrained space). */
int mfgp_gpr_nlml_grad(mfgp_handle* h, const double* X, const double* Y, int N, int d, int P,
                       const double* theta, double noise, double* nlml, double* grad /* [2d+4] */);
/* replaces gpflow GPR.predict_f(Xnew) (tests/test_ho2021_multibin.py:83, tests/test_scipy.py:48):
 * mean [Ns, P], var [Ns] (identical for every output column). */
int mfgp_gpr_predict(mfgp_handle* h, const double* X, const double* Y, int N, int d, int P,
                     const double* Xs, int Ns, const double* theta, double noise,
                     double* mean, double* var);
/* K6: "one GP per k-bin" (gpemulator_singlebin.py:1-14 design; BASELINE config 2): problem
 * b uses y = Y[:, b % ycols] (Y is [N, ycols] row-major with leading dimension ldy),
 * theta[b, :], noise[b]; B > ycols evaluates several hyper-parameter sets (restarts) per
 * bin in one launch.  nlml [B]; grad [B, 2d+4] or NULL; info [B] (int) or NULL.
 * N <= 64 runs one CTA per problem entirely in shared memory. */
int mfgp_gpr_batched_nlml_grad(mfgp_handle* h, const double* X, int N, int d, const double* Y,
                               long ldy, int ycols, int B, const double* theta, const double* noise,
                               double* nlml, double* grad, int* info);

/* Training loop on the device: the Adam branch of MultiFidelityGPModel.optimize (mfgpflow/linear.py:
 * 190-221: loss = -log_marginal_likelihood, tape.gradient w.r.t. the UNCONSTRAINED variables, Keras Adam, noise fixed) for
 * B independent per-bin GPs, nsteps steps without a host round trip.  Per step: theta = softplus(u) -> K6 NLML + gradient
 * -> chain rule (1 - exp(-theta)) -> m, v, u update (ResourceApplyAdam: m += (g-m)(1-beta1), v += (g^2-v)(1-beta2),
 * u -= lr_t m / (sqrt(v) + eps)).  u, m, v are [B, 2d+3] and updated in place (m = v = 0 to start); lr_t [nsteps] holds
 * the per-step factor lr(step) sqrt(1-beta2^t)/(1-beta1^t), computed by the host shim so that the float32-rounded
 * hyper-parameters and the CosineDecay schedule of the reference (quirk Q8) stay in one place.  fix_rho != 0 freezes rho
 * (use_rho=False, linear.py:51-52).  loss_hist [nsteps, B] (or NULL) receives the NLML evaluated BEFORE each update (the
 * reference's loss_history); theta_out [B, 2d+3] (or NULL) the constrained values after the last update.
 * N <= 64 only (the K6 kernel).  Returns >0 if a Cholesky failed at any step (per-problem first failure in info[B]). */
int mfgp_gpr_batched_adam(mfgp_handle* h, const double* X, int N, int d, const double* Y, long ldy, int ycols, int B,
                          double* u, double* m, double* v, const double* noise, const double* lr_t, double beta1,
                          double beta2, double eps, int fix_rho, int nsteps, double* loss_hist, double* theta_out, int* info);

/* ---- Graph-structured multi-fidelity kernel ---------------------------------------------------
 * replaces GraphMultiFidelityKernel.K / K_diag (mfgpflow/graph.py:39-93, :96-115) and, for GraphMultiFidelityGPModel
 * (graph.py:118-188), GPR.log_marginal_likelihood + tape.gradient.  num_lf low-fidelity sources (fidelity column value
 * i in [0, num_lf)) and one high-fidelity level (value num_lf); rows with any other value have zero covariance.
 * gtheta (constrained values), length mfgp_graph_nparams(num_lf, d) = m + m^2 + (m + 1)(d + 1):
 *   [rho (m)] [rho_LF (m x m, row-major, diagonal unused)] [ls_Li (d), var_Li] for i < m, [ls_delta (d), var_delta].
 * Only K(X, X) exists: the reference's rectangular call is not shape-consistent (graph.py:76-79, :91).  K is the FULL
 * N x N matrix including the 1e-6 jitter of graph.py:91; like the reference it is not symmetric once the low-fidelity
 * kernels differ (graph.py:63 uses the row's kernel).  The objective factors the lower triangle of K + noise I and
 * contracts the symmetric sensitivity with the derivative of every entry -- TensorFlow's Cholesky / gradient semantics.
 * grad [nparams + 1] (or NULL): d nlml / d [gtheta, noise]; the entries of the unused rho_LF diagonal are 0. */
int mfgp_graph_nparams(int num_lf, int d);
int mfgp_graph_cov(mfgp_handle* h, const double* X, int N, int d, int num_lf, const double* gtheta, double* K, long ldk);
int mfgp_graph_cov_diag(mfgp_handle* h, const double* X, int N, int d, int num_lf, const double* gtheta, double* out);
int mfgp_graph_gpr_nlml_grad(mfgp_handle* h, const double* X, const double* Y, int N, int d, int P, int num_lf,
                             const double* gtheta, double noise, double* nlml, double* grad);

/* ---- K7/K8: sparse variational GP (whitened, shared inducing points) -------------------
 * replaces gpflow SVGP.elbo / prior_kl / predict_f as called at singlebin_svgp.py:83,97 and
 * linear_svgp.py:177,184,188,199 with kernels SeparateIndependent (W == NULL, L == P,
 * singlebin_svgp.py:39-47) or LinearCoregionalization (W [P, L], linear_svgp.py:121-122). */
typedef struct {
    int L;          /* latent GPs */
    int M;          /* inducing points */
    int P;          /* outputs */
    int B;          /* rows in this (mini)batch */
    int d;          /* input dimension without the fidelity column */
    int hetero;     /* 1: Y is [B, 2P] = [Y_obs | Y_unc], var_eff = lik_var + Y_unc^2 (linear_svgp.py:259) */
    double scale;   /* num_data / B, or 1 (singlebin_svgp.py passes no num_data) */
    double kl_mult; /* loss = -ELBO + (kl_mult - 1) KL   (linear_svgp.py:188) */
    double jitter;  /* gpflow default_jitter() = 1e-6 */
    double lik_lower; /* lower bound of the likelihood-variance transform, used by mfgp_svgp_adam only: gpflow Gaussian
                       * positive(lower=1e-6); HeteroscedasticGaussian positive() = 0 (linear_svgp.py:240) */
    int masked;     /* 1: entries of Y that are NaN are missing outputs and contribute nothing (MaskedGaussian,
                     * notebooks/"demo: missing output.ipynb" cell 2); exclusive with hetero */
    int lik_per_output; /* 1: the likelihood variance is a [P] vector, one per output (same notebook: variance=np.ones(P)) */
} mfgp_svgp_cfg;

/* Forward + hand-derived backward.  Outputs (any grad pointer may be NULL to skip all
 * gradients): elbo, kl (scalars); gradients of loss w.r.t. CONSTRAINED values:
 * gZ [M, d+1] (fidelity column == 0, quirk Q5), gtheta [L, 2d+3], gW [P, L] (NULL if W is
 * NULL), gqmu [M, L], gqsqrt [L, M, M] (lower triangle, rest 0), glik (scalar). */
int mfgp_svgp_elbo_grad(mfgp_handle* h, const mfgp_svgp_cfg* cfg, const double* X, const double* Y,
                        const double* Z, const double* theta, const double* W, const double* q_mu,
                        const double* q_sqrt, double lik_var, double* elbo, double* kl,
                        double* gZ, double* gtheta, double* gW, double* gqmu, double* gqsqrt,
                        double* glik);
/* Same with the likelihood variance passed by pointer (host or device): lik_var and glik hold 1 value, or P values when
 * cfg->lik_per_output is set.  mfgp_svgp_elbo_grad is the scalar-by-value convenience form of this call. */
int mfgp_svgp_elbo_grad_v(mfgp_handle* h, const mfgp_svgp_cfg* cfg, const double* X, const double* Y,
                          const double* Z, const double* theta, const double* W, const double* q_mu,
                          const double* q_sqrt, const double* lik_var, double* elbo, double* kl,
                          double* gZ, double* gtheta, double* gW, double* gqmu, double* gqsqrt,
                          double* glik);
/* mean [Ns, P], var [Ns, P]  (cfg->B is ignored, Ns rows are predicted). */
int mfgp_svgp_predict(mfgp_handle* h, const mfgp_svgp_cfg* cfg, const double* Xs, int Ns,
                      const double* Z, const double* theta, const double* W, const double* q_mu,
                      const double* q_sqrt, double* mean, double* var);

/* Training loop on the device for the SVGP models: the optimize() loops of mfgpflow/singlebin_svgp.py:
 * 64-97 and mfgpflow/linear_svgp.py:153-203 -- full-batch, Keras Adam (+ CosineDecay folded into lr_t, see
 * mfgp_gpr_batched_adam), loss = -ELBO + (kl_mult - 1) KL -- nsteps steps without a host round trip.
 * Flat parameter layout (doubles): [theta L*(2d+3)] [Z M*(d+1)] [W P*L, only if has_W] [q_mu M*L] [q_sqrt L*M*M] [lik_var 1, or P with
 * cfg->lik_per_output].
 * u holds the UNCONSTRAINED values in that layout (theta = softplus(u), lik_var = cfg->lik_lower + softplus(u), the rest identity;
 * the strictly upper part of every q_sqrt matrix must be zero and stays zero); u, m, v are updated in place.
 * mask (n bytes or NULL): 0 freezes an entry (not in trainable_variables).  loss_hist / kl_hist [nsteps] or NULL receive
 * the loss and KL evaluated BEFORE each update (the reference's loss_history / kl_history). */
int mfgp_svgp_adam(mfgp_handle* h, const mfgp_svgp_cfg* cfg, const double* X, const double* Y, int has_W, double* u,
                   double* m, double* v, const unsigned char* mask, const double* lr_t, double beta1, double beta2,
                   double eps, int nsteps, double* loss_hist, double* kl_hist);

/* Data-parallel SVGP training step on device memory only (rows of the minibatch shard across one
 * process per GPU, ONE all-reduce of the flat gradient; reference loops singlebin_svgp.py:79-85, linear_svgp.py:181-190).
 * Per step every rank runs, on the handle's stream and without any host synchronisation:
 *     mfgp_svgp_constrain(u -> c);  mfgp_svgp_elbo_grad_flat(rows of this rank, c -> eg);
 *     all-reduce(sum) of eg [2 + n] in place (the caller's collective: ncclAllReduce on the same stream);
 *     mfgp_svgp_adam_update(u, m, v <- eg).
 * n = mfgp_svgp_flat_size() doubles in the flat layout of mfgp_svgp_adam.  All pointers are DEVICE pointers.
 * eg after mfgp_svgp_elbo_grad_flat: eg[0] = cfg->scale * VE of this rank's rows, eg[1] = KL / nranks, eg[2 + i] = d/d c_i of
 * (-scale VE_rank + kl_mult / nranks KL); after the all-reduce: [scale VE, KL, gradient of -ELBO + (kl_mult - 1) KL].
 * The gradient pieces are written by the kernels straight into eg: no host staging, no concatenation.
 * cfg->B is the number of rows of THIS call; cfg->scale = num_data / (global batch size). */
long mfgp_svgp_flat_size(const mfgp_svgp_cfg* cfg, int has_W);
int mfgp_svgp_constrain(mfgp_handle* h, const mfgp_svgp_cfg* cfg, int has_W, const double* u, double* c);
int mfgp_svgp_elbo_grad_flat(mfgp_handle* h, const mfgp_svgp_cfg* cfg, const double* X, const double* Y, int has_W,
                             const double* c, int nranks, double* eg /* [2 + n] */);
/* Keras-Adam update from the all-reduced eg; lr_t [nsteps] and the step counter *step live on the device (the counter is
 * incremented by the call); loss_hist / kl_hist [nsteps] (device, or NULL) receive entry *step; scratch2: 2 doubles. */
int mfgp_svgp_adam_update(mfgp_handle* h, const mfgp_svgp_cfg* cfg, int has_W, double* u, double* m, double* v,
                          const unsigned char* mask, const double* c, const double* eg, const double* lr_t, int* step,
                          double beta1, double beta2, double eps, double* loss_hist, double* kl_hist, double* scratch2);

/* ---- dense fp64 building blocks (exported for tests / bench / comparators) ------------- */
/* C[m,n] = alpha * op(A) op(B) + beta * C, row-major; transa/transb are 'N' or 'T'.
 * Runs the DMMA (mma.sync m8n8k4 f64) tile kernel; the operand tiles are fed by TMA (cp.async.bulk.tensor + mbarrier), which
 * is why A, B must be 16-byte aligned with even lda / ldb (tensor-map strides are multiples of 16 bytes). */
int mfgp_gemm(mfgp_handle* h, char transa, char transb, int m, int n, int k, double alpha,
              const double* A, long lda, const double* B, long ldb, double beta, double* C, long ldc);
/* In-place lower Cholesky A = L L^T (row-major).  Only the lower triangle is referenced; on return the
 * strictly-upper part of every 128x128 diagonal block is zeroed, upper off-diagonal blocks are untouched. */
int mfgp_potrf(mfgp_handle* h, double* A, int N, long lda);
/* Winv [N, N] = inv(L) for the lower factor produced by mfgp_potrf. */
int mfgp_potrf_inv(mfgp_handle* h, double* A, int N, long lda, double* Winv, long ldw);
/* Y[M, 0:nc] += alpha * A[M, K] X[K, 0:nc] for nc <= 2 right-hand sides (device pointers; A rows 16-byte aligned, K <= 3072).
 * The replicated forward-substitution step of the distributed Cholesky (y[k+1:] -= L[k+1:, k] a_k), HBM-bound. */
int mfgp_tall_skinny_update(mfgp_handle* h, int M, int K, int nc, double alpha, const double* A, long lda, const double* X,
                            long ldx, double* Y, long ldy);

/* One-to-many store: `count` doubles (even, 16-byte aligned) from src into ndst <= MFGP_PEER_MAX destination buffers in ONE
 * kernel on the handle's stream.  The destinations are device pointers valid in this process -- buffers of PEER GPUs mapped
 * over NVLink (symmetric memory, CUDA IPC) or local ones.  Used for the critical-path messages of the distributed Cholesky
 * (dist_chol.py: inv(L_kk) and the early block go to all peers by direct stores instead of an NCCL broadcast); the caller
 * orders the stores against the receivers with its own signal (never synchronises). */
#define MFGP_PEER_MAX 8
int mfgp_peer_store(mfgp_handle* h, const double* src, long count, int ndst, double* const* dsts);

/* Where the temporaries of the following calls on this handle come from.
 *   MFGP_WS_POOL    (default) stream-ordered allocations from the device's default memory pool, per call;
 *   MFGP_WS_MEASURE the same, and the library records the largest number of bytes one call takes;
 *   MFGP_WS_FIXED   one arena of the measured size is allocated NOW (call it outside any stream capture) and every
 *                   following call carves its temporaries out of it, starting again at its beginning (the predecessor's
 *                   temporaries are dead in stream order: keep the calls on one stream).  A call that needs more than was
 *                   measured takes the excess from the pool.
 * Back to MFGP_WS_POOL releases the arena (stream-ordered).  *bytes (may be NULL) receives the measured size (MEASURE ->
 * FIXED or POOL) or 0.  Purpose: library calls captured into a CUDA graph in FIXED mode contain NO allocation nodes, so the
 * graph owns no memory.  Captured in POOL mode the temporaries become graph-owned memory (1.3 GB for the Goku single-bin
 * SVGP step), which the driver keeps reserved after the graph is destroyed until mfgp_graph_mem_trim -- and once trimmed,
 * the next capture pays for reserving it again (measured: 1.0 s instead of 0.07 s per mfgp_svgp_adam call).
 * mfgp_svgp_adam does MEASURE (its eager first step) -> FIXED -> capture -> POOL itself; dist.py does the same around
 * torch.cuda.graph.  Not thread safe per handle, like everything else. */
#define MFGP_WS_POOL 0
#define MFGP_WS_MEASURE 1
#define MFGP_WS_FIXED 2
int mfgp_workspace(mfgp_handle* h, int mode, long* bytes);

/* Give the memory that DESTROYED CUDA graphs of the handle's device still reserve back to the driver
 * (cudaDeviceGraphMemTrim).  Only graphs that captured library calls in MFGP_WS_POOL mode own memory (see mfgp_workspace,
 * which avoids that at its root); call it after destroying such a graph.  Synchronises the stream. */
int mfgp_graph_mem_trim(mfgp_handle* h);

/* FP64 pipe microbenchmarks (bench.py: measured FP64 peak).  kind 0: DFMA, 1: DMMA m8n8k4, 2: both interleaved with equal
 * pipe time (tells whether the two share one datapath: same FLOP/s as either alone, or not: up to twice).
 * Returns achieved FLOP/s in *flops. */
int mfgp_fp64_peak(mfgp_handle* h, int kind, int iters, double* flops);

#ifdef __cplusplus
}
#endif
#endif /* MFGP_H */
