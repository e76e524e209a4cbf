// gpr.cuh -- exact GPR device-side drivers (all pointers are device pointers).
#pragma once
#include "common.cuh"

// batch problems; problem b: RHS = Y[:, c : c + P] with c = ((b_off + b) * per_batch_cols) % ycols (row stride
// ldy; ycols == 0: c = 0), theta_d[b, 2d+3], noise_d[b].  nlml_d [batch]; grad_d [batch, 2d+4] or nullptr.
int gpr_nlml_grad_device(mfgp_handle* h, Scope& sc, const double* X, const double* Y, long ldy, int per_batch_cols,
                         int b_off, int ycols, int N, int d, int P, int batch, const double* theta_d, const double* noise_d,
                         double* nlml_d, double* grad_d, int* info_vec);
int gpr_predict_device(mfgp_handle* h, Scope& sc, const double* X, const double* Y, int N, int d, int P,
                       const double* Xs, int Ns, const double* theta_d, const double* noise_d, double* mean_d,
                       double* var_d);

// Pieces of the pipeline, shared with the graph-kernel model (graph.cu): workspaces of `batch` problems of size N ...
struct GprFactor {
    double *K, *W, *G, *dinv, *logd, *Yw, *a;
    long ld, strideM;
    int Pp;
};
// want_W = false (value only): no W / G buffers (2 N^2 doubles less), and gpr_factor_from_K skips the triangular inverse.
int gpr_factor_alloc(mfgp_handle* h, Scope& sc, int N, int P, int batch, GprFactor& f, bool want_W = true);
// ... K (lower triangle valid, noise included) -> L, W = L^-1, a = W Y; without W: a = L^-1 Y by blocked forward
// substitution over the 128-blocks with the diagonal-block inverses potrf leaves behind (N^2 P flops instead of N^3 / 3) ...
int gpr_factor_from_K(mfgp_handle* h, const double* Y, long ldy, int per_batch_cols, int b_off, int ycols, int N, int P,
                      int batch, int* info_vec, GprFactor& f);
// out <- L^-1 rhs (rhs destroyed) by blocked forward substitution; see gpr.cu.
int gpr_forward_subst(mfgp_handle* h, const GprFactor& f, int N, int batch, double* rhs, double* out, int cols, long ld,
                      long stride);
// ... nlml = 1/2 |a|^2 + P sum log L_ii + N P / 2 log 2 pi ...
void gpr_nlml_from_factor(mfgp_handle* h, const GprFactor& f, int N, int P, int batch, double* nlml_d);
// ... and f.G (lower tiles) <- alpha alpha^T - P K^-1.
int gpr_build_G(mfgp_handle* h, Scope& sc, int N, int P, int batch, GprFactor& f);
