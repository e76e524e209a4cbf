// gpr.cu -- exact multi-fidelity GPR drivers: NLML, analytic gradient, prediction.
//
// Pipeline (all fp64, batched over independent problems):
//   K1 cov (lower tiles + noise)  ->  K2 potrf  ->  K4 trtri: W = L^-1
//   a = W Y, nlml = 1/2 |a|^2 + P sum log L_ii + NP/2 log 2pi          (GPflow logdensities.multivariate_normal)
//   alpha = W^T a,  G = alpha alpha^T - P W^T W  (lower tiles)  ->  K5 contraction with dK/dtheta
// Replaces GPR.log_marginal_likelihood / tape.gradient / predict_f behind
// MultiFidelityGPModel (reference mfgpflow/linear.py:148-156, :203-209).
#include "gpr.cuh"

#include "chol.cuh"
#include "cov.cuh"
#include "gemm.cuh"

namespace {

constexpr double LOG2PI = 1.8378770664093454835606594728112;

// Yw[b][i][p] = Y[i*ldy + col0(b) + p]  (zero padded to Pp columns); col0(b) = ((b_off + b) * per_batch_cols) % ycols
__global__ void pack_rhs_kernel(const double* __restrict__ Y, long ldy, int N, int P, int Pp, int per_batch_cols,
                                int b_off, int ycols, double* __restrict__ Yw) {
    const int b = blockIdx.y;
    const long col0 = ycols > 0 ? ((long)(b_off + b) * per_batch_cols) % ycols : 0;
    const long tot = (long)N * Pp;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < tot; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / Pp), pcol = (int)(idx % Pp);
        Yw[(long)b * tot + idx] = pcol < P ? Y[(long)i * ldy + col0 + pcol] : 0.0;
    }
}

__device__ inline double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    __syncthreads();
    return t;  // valid on thread 0
}

__global__ void nlml_kernel(const double* __restrict__ a, int N, int Pp, int P, const double* __restrict__ logd,
                            double* __restrict__ out) {
    __shared__ double sh[8];
    const int b = blockIdx.x;
    double q = 0.0, ld = 0.0;
    const double* ab = a + (long)b * N * Pp;
    for (long i = threadIdx.x; i < (long)N * Pp; i += blockDim.x) q = fma(ab[i], ab[i], q);
    for (int i = threadIdx.x; i < N; i += blockDim.x) ld += logd[(long)b * N + i];
    q = block_sum(q, sh);
    ld = block_sum(ld, sh);
    if (threadIdx.x == 0) out[b] = 0.5 * q + P * ld + 0.5 * (double)N * P * LOG2PI;
}

// var[j] = kss[j] - sum_i As[i][j]^2
__global__ void predict_var_kernel(const double* __restrict__ As, int N, int Ns, long ld, const double* __restrict__ kss,
                                   double* __restrict__ var) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Ns) return;
    double s = 0.0;
    for (int i = 0; i < N; ++i) {
        const double v = As[(long)i * ld + j];
        s = fma(v, v, s);
    }
    var[j] = kss[j] - s;
}

using Factor = GprFactor;

// Assemble + factor + invert + a = W Y for `batch` problems.  theta_d/noise_d are device arrays.
int factor(mfgp_handle* h, Scope& sc, const double* X, const double* Y, long ldy, int per_batch_cols, int b_off,
           int ycols, int N, int d, int P, int batch, const double* theta_d, const double* noise_d, int* info_vec, Factor& f,
           bool want_W = true) {
    MFGP_TRY(gpr_factor_alloc(h, sc, N, P, batch, f, want_W));
    CovArgs c{};
    c.Xa = X; c.Na = N; c.Xb = X; c.Nb = N; c.d = d;
    c.theta = theta_d; c.theta_stride = 2 * d + 3;
    c.K = f.K; c.ldk = f.ld; c.strideK = f.strideM;
    c.symmetric = 1; c.mirror = 0;
    c.diag_add = 0.0; c.diag_add_vec = noise_d;
    c.batch = batch;
    if (launch_cov(h->stream, c)) return mfgp_fail(h, MFGP_ERR_CUDA, "cov launch failed");
    return gpr_factor_from_K(h, Y, ldy, per_batch_cols, b_off, ycols, N, P, batch, info_vec, f);
}

}  // namespace

int gpr_factor_alloc(mfgp_handle* h, Scope& sc, int N, int P, int batch, GprFactor& f, bool want_W) {
    f.ld = round_up(N, 2);
    f.strideM = (long)N * f.ld;
    f.Pp = (int)round_up(P, 2);
    f.K = sc.alloc<double>((size_t)batch * f.strideM);
    f.W = want_W ? sc.alloc<double>((size_t)batch * f.strideM) : nullptr;
    f.G = want_W ? sc.alloc<double>((size_t)batch * f.strideM) : nullptr;  // trtri scratch, then G
    f.dinv = sc.alloc<double>((size_t)chol_dinv_count(N, batch));
    f.logd = sc.alloc<double>((size_t)batch * N);
    f.Yw = sc.alloc<double>((size_t)batch * N * f.Pp);
    f.a = sc.alloc<double>((size_t)batch * N * f.Pp);
    (void)h;
    return sc.ok ? 0 : MFGP_ERR_CUDA;
}

// out[batch][N, ld] (cols columns) <- L^-1 rhs, block row by block row: out_k = inv(L_kk) rhs_k with the diagonal-block
// inverses potrf left in f.dinv, then rhs_{k+1:} -= L_{k+1:, k} out_k.  rhs is destroyed; L is read once (4.3 GB at
// N = 32 768).  ld even, both arrays 16-byte aligned (the TMA-fed GEMM's contract).
int gpr_forward_subst(mfgp_handle* h, const GprFactor& f, int N, int batch, double* rhs, double* out, int cols, long ld,
                      long stride) {
    cudaStream_t s = h->stream;
    const long strideD = (long)chol_nblk(N) * CHOL_NB * CHOL_NB;
    for (int k0 = 0; k0 < N; k0 += CHOL_NB) {
        const int nb = N - k0 < CHOL_NB ? N - k0 : CHOL_NB, k1 = k0 + nb;
        GemmArgs t;  // out_k = inv(L_kk) rhs_k
        t.M = nb; t.N = cols; t.K = nb;
        t.A = f.dinv + (long)(k0 / CHOL_NB) * CHOL_NB * CHOL_NB; t.lda = CHOL_NB; t.strideA = strideD;
        t.B = rhs + (long)k0 * ld; t.ldb = ld; t.strideB = stride;
        t.C = out + (long)k0 * ld; t.ldc = ld; t.strideC = stride;
        t.batch = batch;
        if (launch_gemm(s, t)) return mfgp_fail(h, MFGP_ERR_CUDA, "gemm (inv(L_kk) rhs_k) failed");
        if (k1 >= N) break;
        GemmArgs u;  // rhs_{k+1:} -= L_{k+1:, k} out_k
        u.M = N - k1; u.N = cols; u.K = nb;
        u.alpha = -1.0; u.beta = 1.0;
        u.A = f.K + (long)k1 * f.ld + k0; u.lda = f.ld; u.strideA = f.strideM;
        u.B = out + (long)k0 * ld; u.ldb = ld; u.strideB = stride;
        u.C = rhs + (long)k1 * ld; u.ldc = ld; u.strideC = stride;
        u.batch = batch;
        if (launch_gemm(s, u)) return mfgp_fail(h, MFGP_ERR_CUDA, "gemm (forward substitution update) failed");
    }
    return 0;
}

// f.K holds the (lower triangle of the) noisy covariance: potrf -> W = L^-1 -> a = W Y.
int gpr_factor_from_K(mfgp_handle* h, const double* Y, long ldy, int per_batch_cols, int b_off, int ycols, int N, int P,
                      int batch, int* info_vec, GprFactor& f) {
    cudaStream_t s = h->stream;
    CholArgs ch{};
    ch.A = f.K; ch.N = N; ch.lda = f.ld; ch.strideA = f.strideM; ch.batch = batch;
    ch.dinv = f.dinv; ch.logd = f.logd; ch.d_info = h->d_info; ch.aux = h->aux_stream; ch.ev = h->ev; ch.info_vec = info_vec;
    if (launch_potrf(s, ch)) return mfgp_fail(h, MFGP_ERR_CUDA, "potrf launch failed");
    pack_rhs_kernel<<<dim3(64, batch), 256, 0, s>>>(Y, ldy, N, P, f.Pp, per_batch_cols, b_off, ycols, f.Yw);
    if (!f.W)  // value only: a = L^-1 Yw without the N^3 / 3 inverse and without the W and G buffers
        return gpr_forward_subst(h, f, N, batch, f.Yw, f.a, f.Pp, f.Pp, (long)N * f.Pp);
    if (launch_trtri(s, ch, f.W, f.ld, f.strideM, f.G)) return mfgp_fail(h, MFGP_ERR_CUDA, "trtri launch failed");

    GemmArgs g;  // a = W Yw
    g.transA = false; g.transB = false;
    g.M = N; g.N = f.Pp; g.K = N;
    g.A = f.W; g.lda = f.ld; g.strideA = f.strideM;
    g.B = f.Yw; g.ldb = f.Pp; g.strideB = (long)N * f.Pp;
    g.C = f.a; g.ldc = f.Pp; g.strideC = (long)N * f.Pp;
    g.batch = batch;
    g.krange = KR_HI_I;
    if (launch_gemm(s, g)) return mfgp_fail(h, MFGP_ERR_CUDA, "gemm (a = W Y) failed");
    return 0;
}

void gpr_nlml_from_factor(mfgp_handle* h, const GprFactor& f, int N, int P, int batch, double* nlml_d) {
    nlml_kernel<<<batch, 256, 0, h->stream>>>(f.a, N, f.Pp, P, f.logd, nlml_d);
}

// f.G (lower tiles) <- alpha alpha^T - P K^-1 with alpha = W^T a, K^-1 = W^T W
int gpr_build_G(mfgp_handle* h, Scope& sc, int N, int P, int batch, GprFactor& f) {
    cudaStream_t s = h->stream;
    const long strideV = (long)N * f.Pp;
    double* alpha = sc.alloc<double>((size_t)batch * strideV);
    if (!sc.ok) return MFGP_ERR_CUDA;
    GemmArgs g;  // alpha = W^T a
    g.transA = true; g.transB = false;
    g.M = N; g.N = f.Pp; g.K = N;
    g.A = f.W; g.lda = f.ld; g.strideA = f.strideM;
    g.B = f.a; g.ldb = f.Pp; g.strideB = strideV;
    g.C = alpha; g.ldc = f.Pp; g.strideC = strideV;
    g.batch = batch;
    g.krange = KR_LO_I;
    if (launch_gemm(s, g)) return mfgp_fail(h, MFGP_ERR_CUDA, "gemm (alpha) failed");

    GemmArgs o;  // G = alpha alpha^T  (lower tiles)
    o.transA = false; o.transB = true;
    o.M = N; o.N = N; o.K = f.Pp;
    o.A = alpha; o.lda = f.Pp; o.strideA = strideV;
    o.B = alpha; o.ldb = f.Pp; o.strideB = strideV;
    o.C = f.G; o.ldc = f.ld; o.strideC = f.strideM;
    o.batch = batch;
    o.lower_only = 1;
    if (launch_gemm(s, o)) return mfgp_fail(h, MFGP_ERR_CUDA, "gemm (alpha alpha^T) failed");

    GemmArgs k;  // G -= P * W^T W   (K^-1 = L^-T L^-1)
    k.transA = true; k.transB = false;
    k.M = N; k.N = N; k.K = N;
    k.alpha = -(double)P; k.beta = 1.0;
    k.A = f.W; k.lda = f.ld; k.strideA = f.strideM;
    k.B = f.W; k.ldb = f.ld; k.strideB = f.strideM;
    k.C = f.G; k.ldc = f.ld; k.strideC = f.strideM;
    k.batch = batch;
    k.krange = KR_LO_MAXIJ;
    k.lower_only = 1;
    if (launch_gemm(s, k)) return mfgp_fail(h, MFGP_ERR_CUDA, "gemm (W^T W) failed");
    return 0;
}

int gpr_nlml_grad_device(mfgp_handle* h, Scope& sc, const double* X, const double* Y, long ldy, int per_batch_cols,
                         int b_off, int ycols, int N, int d, int P, int batch, const double* theta_d, const double* noise_d,
                         double* nlml_d, double* grad_d, int* info_vec) {
    cudaStream_t s = h->stream;
    Factor f;
    MFGP_TRY(factor(h, sc, X, Y, ldy, per_batch_cols, b_off, ycols, N, d, P, batch, theta_d, noise_d, info_vec, f, grad_d != nullptr));
    gpr_nlml_from_factor(h, f, N, P, batch, nlml_d);
    if (!grad_d) return 0;
    MFGP_TRY(gpr_build_G(h, sc, N, P, batch, f));

    CovGradArgs cg{};
    cg.Xa = X; cg.Na = N; cg.Xb = X; cg.Nb = N; cg.d = d;
    cg.theta = theta_d; cg.theta_stride = 2 * d + 3;
    cg.G = f.G; cg.ldg = f.ld; cg.strideG = f.strideM;
    cg.sym_lower = 1;
    cg.out = grad_d; cg.out_stride = 2 * d + 4;
    cg.out_scale = -0.5;  // d(nlml)/dtheta = -1/2 sum_ij G_ij dK_ij/dtheta
    cg.accumulate = 0;
    cg.rowgrad = nullptr;
    cg.batch = batch;
    cg.partial = sc.alloc<double>((size_t)cov_grad_partial_count(cg));
    if (!sc.ok) return MFGP_ERR_CUDA;
    if (launch_cov_grad(s, cg)) return mfgp_fail(h, MFGP_ERR_CUDA, "cov_grad launch failed");
    return 0;
}

int gpr_predict_device(mfgp_handle* h, Scope& sc, const double* X, const double* Y, int N, int d, int P,
                       const double* Xs, int Ns, const double* theta_d, const double* noise_d, double* mean_d,
                       double* var_d) {
    cudaStream_t s = h->stream;
    Factor f;
    // A_s = L^-1 K_s by forward substitution (N^2 Ns flops in k = 128 updates) beats building W = L^-1 first (N^3 / 3 more
    // flops) for as long as Ns < N: measured at N = 16 384 (scripts/predict_once.py) 61 + 0.0098 Ns ms against
    // 100 + 0.0077 Ns ms for W + one full-rate GEMM.
    const bool want_W = Ns > N;
    MFGP_TRY(factor(h, sc, X, Y, P, 0, 0, 0, N, d, P, 1, theta_d, noise_d, nullptr, f, want_W));
    const long lds = round_up(Ns, 2);
    double* Ks = sc.alloc<double>((size_t)N * lds);
    double* As = sc.alloc<double>((size_t)N * lds);
    double* kss = sc.alloc<double>((size_t)Ns);
    if (!sc.ok) return MFGP_ERR_CUDA;
    CovArgs c{};
    c.Xa = X; c.Na = N; c.Xb = Xs; c.Nb = Ns; c.d = d;
    c.theta = theta_d; c.theta_stride = 2 * d + 3;
    c.K = Ks; c.ldk = lds; c.strideK = 0;
    c.batch = 1;
    if (launch_cov(s, c)) return mfgp_fail(h, MFGP_ERR_CUDA, "cov launch failed");
    if (launch_cov_diag(s, Xs, Ns, d, theta_d, 0, kss, 0, 1)) return mfgp_fail(h, MFGP_ERR_CUDA, "cov_diag failed");
    if (want_W) {
        GemmArgs g;  // As = W Ks
        g.transA = false; g.transB = false;
        g.M = N; g.N = Ns; g.K = N;
        g.A = f.W; g.lda = f.ld;
        g.B = Ks; g.ldb = lds;
        g.C = As; g.ldc = lds;
        g.krange = KR_HI_I;
        if (launch_gemm(s, g)) return mfgp_fail(h, MFGP_ERR_CUDA, "gemm (As) failed");
    } else {
        MFGP_TRY(gpr_forward_subst(h, f, N, 1, Ks, As, Ns, lds, 0));
    }
    predict_var_kernel<<<(Ns + 127) / 128, 128, 0, s>>>(As, N, Ns, lds, kss, var_d);
    GemmArgs m;  // mean = As^T a
    m.transA = true; m.transB = false;
    m.M = Ns; m.N = P; m.K = N;
    m.A = As; m.lda = lds;
    m.B = f.a; m.ldb = f.Pp;
    m.C = mean_d; m.ldc = P;
    if (launch_gemm(s, m)) return mfgp_fail(h, MFGP_ERR_CUDA, "gemm (mean) failed");
    return 0;
}
