// chol.cuh -- K2/K4: blocked fp64 Cholesky and triangular inverse built on the DMMA GEMM.
#pragma once
#include <cuda_runtime.h>

#define CHOL_NB 128

struct CholArgs {
    double* A;  // [batch][N, lda] row-major, lower triangle valid on entry; L on exit
    int N;
    long lda, strideA;
    int batch;
    double* dinv;   // workspace [batch][nblk][128*128]: inverses of the diagonal blocks of L
    double* logd;   // [batch][N]: log(L_ii)
    int* d_info;    // first non-positive pivot (1-based global index), CAS'ed from 0
    int* info_vec;  // optional [batch]
    cudaStream_t aux = nullptr;     // optional look-ahead stream for the diagonal-block / panel chain
    cudaEvent_t* ev = nullptr;      // 4 events (disable-timing) when aux is set
};
inline int chol_nblk(int N) { return (N + CHOL_NB - 1) / CHOL_NB; }
inline long chol_dinv_count(int N, int batch) { return (long)batch * chol_nblk(N) * CHOL_NB * CHOL_NB; }

// A <- L (lower, strict upper of the diagonal blocks zeroed), dinv, logd.
int launch_potrf(cudaStream_t s, const CholArgs& a);

// W [batch][N, ldw] <- inv(L) (lower, everything above the diagonal zero) from L and dinv.
// scratch: [batch][N, ldw] doubles.
int launch_trtri(cudaStream_t s, const CholArgs& a, double* W, long ldw, long strideW, double* scratch);
