// gemm.cuh -- batched fp64 GEMM on the DMMA (mma.sync m8n8k4 f64) path.
#pragma once
#include <cuda_runtime.h>

// Row-major:  C[M,N] = alpha * op(A) * op(B) + beta * C
//   transA == false: A is [M,K] (A[i*lda + k]);  true: A is [K,M] (A[k*lda + i])
//   transB == false: B is [K,N] (B[k*ldb + j]);  true: B is [N,K] (B[j*ldb + k])
// Requirements: A, B 16-byte aligned, lda/ldb/strides even (library workspaces are padded).
//
// Triangular structure is expressed as a per-tile contraction range so that blocks that are
// known to be zero are never loaded:   k in [klo(tile), khi(tile))
enum GemmKRange {
    KR_FULL = 0,
    KR_LO_I = 1,      // k >= i0            (e.g. A upper-triangular in (i,k))
    KR_LO_J = 2,      // k >= j0            (B lower-triangular as [K,N]: B[k,j] = 0 for k < j)
    KR_LO_MAXIJ = 3,  // k >= max(i0, j0)
    KR_HI_I = 4,      // k <  i0 + BM       (A lower-triangular: A[i,k] = 0 for k > i)
    KR_HI_J = 8,      // k <  j0 + BN
    KR_HI_MINIJ = 12  // k <  min(i0+BM, j0+BN)
};

struct GemmArgs {
    bool transA = false, transB = false;
    int M = 0, N = 0, K = 0;
    double alpha = 1.0, beta = 0.0;
    const double* A = nullptr;
    long lda = 0, strideA = 0;
    const double* B = nullptr;
    long ldb = 0, strideB = 0;
    double* C = nullptr;
    long ldc = 0, strideC = 0;
    int batch = 1;
    int batch2 = 1;  // outer batch: blockIdx.z = b1 + batch * b2, pointer += b1*stride + b2*stride2
    long strideA2 = 0, strideB2 = 0, strideC2 = 0;
    int krange = KR_FULL;  // OR of one KR_LO_* and one KR_HI_*
    int lower_only = 0;    // skip tiles strictly above the diagonal (symmetric / triangular outputs)
    int small_tiles = -1;  // tile config: -1 auto, 0 = 128x128, 1 = 64x64, 2 = 128x64
};

int launch_gemm(cudaStream_t s, const GemmArgs& a);
