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
