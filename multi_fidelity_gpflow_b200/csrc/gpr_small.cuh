// gpr_small.cuh -- batched small-matrix GPR NLML+grad (K6: one warp per problem, persistent grid).
#pragma once
#include <cuda_runtime.h>

#define MFGP_SMALL_MAX_N 64
#define MFGP_SMALL_MAX_D 16

struct SmallArgs {
    const double* X;  // [N, d+1] shared by all problems
    int N, d;
    const double* Y;  // [N, ldy], problem b uses column b % ycols
    long ldy;
    int ycols;
    int B;
    const double* theta;  // [B, 2d+3]
    const double* noise;  // [B]
    double* nlml;         // [B]
    double* grad;         // [B, 2d+4] or nullptr
    int* info;            // [B] or nullptr
    int* d_info;          // handle-wide first failure
    int NP;               // filled by the launcher
    int prob0;            // v4: index of this launch's first problem in the caller's batch (selects the Y column)
    double* scratch;      // v4: K^L of every problem in flight (L2-resident), filled by the launcher
};
int launch_gpr_small_v4(cudaStream_t s, const SmallArgs& a);  // persistent grid, one warp per problem, 12 warps per SM
inline int launch_gpr_small(cudaStream_t s, const SmallArgs& a) { return launch_gpr_small_v4(s, a); }
