// graph.cu -- GraphMultiFidelityKernel (reference mfgpflow/graph.py:7-115) and the exact-GP training objective of
// GraphMultiFidelityGPModel (graph.py:118-188): m low-fidelity sources with learnable cross-correlations plus one
// high-fidelity level,  f_H = sum_i rho_i f_Li + delta.   SURVEY 8(f) rank 3.
//
// Fidelity column: i in [0, m) = low-fidelity source i, m = high fidelity; any other value (NaN included) gives an all-zero
// row / column (graph.py:54 starts from zeros and scatters only the listed blocks), exact float equality like graph.py:45-48.
//   K[a in Li, b in Lj] = (i == j ? 1 : rho_LF[i, j]) k_Li(x_a, x_b)        (graph.py:57-66: the ROW's kernel -- the matrix is
//                                                                             not symmetric once k_Li != k_Lj or rho_LF is not)
//   K[a in Li, b in H ] = rho_i k_Li(x_a, x_b),   K[a in H, b in Li] = rho_i k_Li(x_a, x_b)              (graph.py:69-79)
//   K[a in H,  b in H ] = sum_i rho_i^2 k_Li(x_a, x_b) + k_delta(x_a, x_b)                                 (graph.py:82-88)
//   K += 1e-6 I                                                                                            (graph.py:91)
// Only K(X, X) is well defined in the reference (the rectangular scatter indices and the eye(N) jitter are shape-consistent
// for X2 = X only), so this file implements the symmetric call, K_diag (graph.py:96-115) and the GPR objective.
//
// Objective semantics reproduced from TensorFlow (which the reference differentiates through): tf.linalg.cholesky reads the
// LOWER triangle of K + noise I, and its registered gradient returns the SYMMETRISED sensitivity 1/2 (S + S^T); the
// chain rule then runs over ALL N^2 entries of K.  So  d nlml / d theta = -1/2 sum_{a,b} G_ab dK_ab / d theta  with the
// symmetric G = alpha alpha^T - P K_n^-1 built from the lower-triangle factor, and dK_ab the derivative of the entry as
// written above (both triangles, asymmetric blocks included).
//
// Parameter vector gtheta (CONSTRAINED values), length m + m^2 + (m + 1)(d + 1):
//   [rho_0 .. rho_{m-1}] [rho_LF row-major m x m (diagonal unused)] [ls_L0 (d), var_L0] ... [ls_L{m-1} (d), var_L{m-1}] [ls_delta (d), var_delta]
#include <cuda_runtime.h>

#include "common.cuh"
#include "cov.cuh"
#include "gpr.cuh"

#define MFGP_GRAPH_MAX_LF 4
#define MFGP_GRAPH_MAX_PARAMS (MFGP_GRAPH_MAX_LF + MFGP_GRAPH_MAX_LF * MFGP_GRAPH_MAX_LF + (MFGP_GRAPH_MAX_LF + 1) * (MFGP_MAX_D + 1))

namespace {

constexpr double GRAPH_JITTER = 1e-6;  // graph.py:91

__host__ __device__ inline int graph_nparams(int m, int d) { return m + m * m + (m + 1) * (d + 1); }
__device__ __forceinline__ int off_kernel(int i, int m, int d) { return m + m * m + i * (d + 1); }  // i == m: delta

__device__ __forceinline__ int fid_class(double f, int m) {
    for (int i = 0; i <= m; ++i)
        if (f == (double)i) return i;
    return -1;
}

// exp(-1/2 sum ((xa - xb) / ls)^2), without the variance
__device__ __forceinline__ double se_shape(const double* __restrict__ xa, const double* __restrict__ xb,
                                           const double* __restrict__ ls, int d) {
    double e = 0.0;
    for (int q = 0; q < d; ++q) {
        const double t = (xa[q] - xb[q]) / ls[q];
        e = fma(t, t, e);
    }
    return exp(-0.5 * e);
}

__global__ void graph_cov_kernel(const double* __restrict__ X, int N, int d, int m, const double* __restrict__ gth,
                                 const double* __restrict__ noise, double* __restrict__ K, long ld) {
    const int a = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= N || b >= N) return;
    const double *xa = X + (long)a * (d + 1), *xb = X + (long)b * (d + 1);
    const int ca = fid_class(xa[d], m), cb = fid_class(xb[d], m);
    double v = 0.0;
    if (ca >= 0 && cb >= 0) {
        if (ca < m && cb < m) {
            const double* kp = gth + off_kernel(ca, m, d);
            v = (ca == cb ? 1.0 : gth[m + ca * m + cb]) * kp[d] * se_shape(xa, xb, kp, d);
        } else if (ca < m || cb < m) {
            const int i = ca < m ? ca : cb;
            const double* kp = gth + off_kernel(i, m, d);
            v = gth[i] * kp[d] * se_shape(xa, xb, kp, d);
        } else {
            for (int i = 0; i < m; ++i) {
                const double* kp = gth + off_kernel(i, m, d);
                v = fma(gth[i] * gth[i], kp[d] * se_shape(xa, xb, kp, d), v);
            }
            const double* kp = gth + off_kernel(m, m, d);
            v += kp[d] * se_shape(xa, xb, kp, d);
        }
    }
    if (a == b) v += GRAPH_JITTER + (noise ? noise[0] : 0.0);
    K[(long)a * ld + b] = v;
}

__global__ void graph_cov_diag_kernel(const double* __restrict__ X, int N, int d, int m, const double* __restrict__ gth,
                                      double* __restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= N) return;
    const int c = fid_class(X[(long)a * (d + 1) + d], m);
    double v = 0.0;  // graph.py:101 zeros; no jitter on K_diag
    if (c >= 0 && c < m) v = gth[off_kernel(c, m, d) + d];
    else if (c == m) {
        for (int i = 0; i < m; ++i) v = fma(gth[i] * gth[i], gth[off_kernel(i, m, d) + d], v);
        v += gth[off_kernel(m, m, d) + d];
    }
    out[a] = v;
}

// One warp per row a of the FULL matrix; lanes stride over the columns, every lane keeps a private accumulator per
// parameter (local memory, L1-resident), then a fixed-order warp reduction writes partial[a][0 .. np] (np = trace slot).
__global__ void graph_cov_grad_kernel(const double* __restrict__ X, int N, int d, int m, const double* __restrict__ gth,
                                      const double* __restrict__ G, long ldg, double* __restrict__ partial) {
    const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (a >= N) return;
    const int np = graph_nparams(m, d);
    double acc[MFGP_GRAPH_MAX_PARAMS + 1];
    for (int q = 0; q <= np; ++q) acc[q] = 0.0;
    const double* xa = X + (long)a * (d + 1);
    const int ca = fid_class(xa[d], m);
    if (ca >= 0) {
        for (int b = lane; b < N; b += 32) {
            const double* xb = X + (long)b * (d + 1);
            const int cb = fid_class(xb[d], m);
            if (cb < 0) continue;
            const double g = a >= b ? G[(long)a * ldg + b] : G[(long)b * ldg + a];  // symmetric G, lower triangle stored
            // contribution of kernel i with coefficient coef: K_ab has the term coef * var_i * shape_i
            auto kernel_terms = [&](int i, double coef) {
                const int o = off_kernel(i, m, d);
                const double* kp = gth + o;
                const double sh = se_shape(xa, xb, kp, d);
                const double t = g * coef * sh;  // d/d var
                acc[o + d] += t;
                const double tv = t * kp[d];
                for (int q = 0; q < d; ++q) {
                    const double dx = xa[q] - xb[q];
                    acc[o + q] = fma(tv * dx * dx, 1.0 / (kp[q] * kp[q] * kp[q]), acc[o + q]);
                }
                return kp[d] * sh;  // k_i(x_a, x_b)
            };
            if (ca < m && cb < m) {
                const double coef = ca == cb ? 1.0 : gth[m + ca * m + cb];
                const double k = kernel_terms(ca, coef);
                if (ca != cb) acc[m + ca * m + cb] = fma(g, k, acc[m + ca * m + cb]);
            } else if (ca < m || cb < m) {
                const int i = ca < m ? ca : cb;
                const double k = kernel_terms(i, gth[i]);
                acc[i] = fma(g, k, acc[i]);
            } else {
                for (int i = 0; i < m; ++i) {
                    const double k = kernel_terms(i, gth[i] * gth[i]);
                    acc[i] = fma(2.0 * gth[i] * g, k, acc[i]);
                }
                kernel_terms(m, 1.0);
            }
        }
    }
    if (lane == 0) acc[np] = G[(long)a * ldg + a];  // noise (and jitter) sit on the whole diagonal, dead rows included
    for (int q = 0; q <= np; ++q) {
        double v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) partial[(long)a * (np + 1) + q] = v;
    }
}

__global__ void graph_grad_reduce_kernel(const double* __restrict__ partial, int N, int np1, double scale, double* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= np1) return;
    double s = 0.0;
    for (int a = 0; a < N; ++a) s += partial[(long)a * np1 + q];  // rows in index order: deterministic
    out[q] = scale * s;
}

int graph_check(mfgp_handle* h, const void* X, const void* gth, int N, int d, int m, const char* who) {
    if (!X || !gth || N < 1 || d < 1 || d > MFGP_MAX_D || m < 1 || m > MFGP_GRAPH_MAX_LF)
        return mfgp_fail(h, MFGP_ERR_ARG, "%s: bad argument (1 <= num_LF <= %d, 1 <= d <= %d)", who, MFGP_GRAPH_MAX_LF, MFGP_MAX_D);
    return 0;
}

}  // namespace

extern "C" {

int mfgp_graph_nparams(int num_lf, int d) { return graph_nparams(num_lf, d); }

int mfgp_graph_cov(mfgp_handle* h, const double* X, int N, int d, int num_lf, const double* gtheta, double* K, long ldk) {
    if (!h) return MFGP_ERR_ARG;
    MFGP_TRY(graph_check(h, X, gtheta, N, d, num_lf, "mfgp_graph_cov"));
    if (!K || ldk < N) return mfgp_fail(h, MFGP_ERR_ARG, "mfgp_graph_cov: bad K / ldk");
    cudaSetDevice(h->device);
    Scope sc(h);
    const double* dX = sc.in(X, (size_t)N * (d + 1));
    const double* dth = sc.in(gtheta, graph_nparams(num_lf, d));
    double* dK = sc.out(K, (size_t)N * ldk);
    if (!sc.ok) return sc.finish();
    dim3 blk(16, 16), grid((N + 15) / 16, (N + 15) / 16);
    graph_cov_kernel<<<grid, blk, 0, h->stream>>>(dX, N, d, num_lf, dth, nullptr, dK, ldk);
    return sc.finish();
}

int mfgp_graph_cov_diag(mfgp_handle* h, const double* X, int N, int d, int num_lf, const double* gtheta, double* out) {
    if (!h) return MFGP_ERR_ARG;
    MFGP_TRY(graph_check(h, X, gtheta, N, d, num_lf, "mfgp_graph_cov_diag"));
    if (!out) return mfgp_fail(h, MFGP_ERR_ARG, "mfgp_graph_cov_diag: out is NULL");
    cudaSetDevice(h->device);
    Scope sc(h);
    const double* dX = sc.in(X, (size_t)N * (d + 1));
    const double* dth = sc.in(gtheta, graph_nparams(num_lf, d));
    double* dout = sc.out(out, N);
    if (!sc.ok) return sc.finish();
    graph_cov_diag_kernel<<<(N + 255) / 256, 256, 0, h->stream>>>(dX, N, d, num_lf, dth, dout);
    return sc.finish();
}

int mfgp_graph_gpr_nlml_grad(mfgp_handle* h, const double* X, const double* Y, int N, int d, int P, int num_lf,
                             const double* gtheta, double noise, double* nlml, double* grad) {
    if (!h) return MFGP_ERR_ARG;
    MFGP_TRY(graph_check(h, X, gtheta, N, d, num_lf, "mfgp_graph_gpr_nlml_grad"));
    if (!Y || !nlml || P < 1) return mfgp_fail(h, MFGP_ERR_ARG, "mfgp_graph_gpr_nlml_grad: bad argument");
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    const int np = graph_nparams(num_lf, d);
    Scope sc(h);
    const double* dX = sc.in(X, (size_t)N * (d + 1));
    const double* dY = sc.in(Y, (size_t)N * P);
    const double* dth = sc.in(gtheta, np);
    const double* dnz = sc.in(&noise, 1);
    double* dn = sc.out(nlml, 1);
    double* dg = grad ? sc.out(grad, np + 1) : nullptr;
    if (!sc.ok) return sc.finish();
    GprFactor f;
    MFGP_TRY(gpr_factor_alloc(h, sc, N, P, 1, f));
    dim3 blk(16, 16), grid((N + 15) / 16, (N + 15) / 16);
    graph_cov_kernel<<<grid, blk, 0, s>>>(dX, N, d, num_lf, dth, dnz, f.K, f.ld);  // potrf reads the lower triangle only
    MFGP_TRY(gpr_factor_from_K(h, dY, P, 0, 0, 0, N, P, 1, nullptr, f));
    gpr_nlml_from_factor(h, f, N, P, 1, dn);
    if (dg) {
        MFGP_TRY(gpr_build_G(h, sc, N, P, 1, f));
        double* partial = sc.alloc<double>((size_t)N * (np + 1));
        if (!sc.ok) return sc.finish();
        graph_cov_grad_kernel<<<(N * 32 + 127) / 128, 128, 0, s>>>(dX, N, d, num_lf, dth, f.G, f.ld, partial);
        graph_grad_reduce_kernel<<<(np + 1 + 63) / 64, 64, 0, s>>>(partial, N, np + 1, -0.5, dg);
    }
    return sc.finish();
}

}  // extern "C"
