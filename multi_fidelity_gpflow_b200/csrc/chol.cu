// chol.cu -- K2: blocked right-looking fp64 Cholesky (lower, row-major); K4: triangular inverse.
//
// potrf (NB = 128):  for each diagonal block
//     diag kernel   one CTA factors the 128x128 block with the trailing matrix held in
//                   REGISTERS (2-D cyclic ownership, one published column + one barrier per
//                   step) and inverts the factor in shared memory            [latency bound]
//     panel         A21 <- A21 * inv(L11)^T      (DMMA GEMM, in place)
//     trailing      A22 <- A22 - A21 A21^T       (DMMA GEMM, lower tiles only) [FP64 pipe bound]
// trtri: inv(L) by log2(N/128) levels of batched merges
//     [W11 0; W21 W22],  W21 = -W22 (L21 W11)    two DMMA GEMMs per level, every level a full grid.
// These replace tf.linalg.cholesky / triangular_solve behind GPflow's GPR / SVGP objectives
// (reference call sites linear.py:206, singlebin_svgp.py:83, linear_svgp.py:184).
#include "chol.cuh"

#include "gemm.cuh"

namespace {

constexpr int NB = CHOL_NB;
constexpr int SP = NB + 1;  // odd pitch: conflict-free column and row walks

struct DiagArgs {
    double* A;  // top-left of the diagonal block (batch 0)
    long lda, strideA;
    int nb;       // rows in this block (<= 128)
    int k0;       // global index of the first row (for info)
    double* dinv;  // [128*128] for this block (batch 0)
    long stride_dinv;
    double* logd;  // + k0 (batch 0)
    long stride_logd;
    int* d_info;
    int* info_vec;
};

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(DiagArgs p) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                 // [128][129]: lower = L, strict upper = inv(L)^T
    double* col = S + NB * SP;      // [2][128]
    double* invd = col + 2 * NB;    // [128]
    double* piv = invd + NB;        // [128]
    __shared__ int bad;
    const int nb = p.nb;
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tk = tid & 15;
    const long bz = blockIdx.x;
    double* __restrict__ A = p.A + bz * p.strideA;
    if (tid == 0) bad = 0;

    double r[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int i = ti + 16 * a, k = tk + 16 * b;
            r[a][b] = (k <= i && i < nb) ? A[(long)i * p.lda + k] : 0.0;
        }
    __syncthreads();

#pragma unroll
    for (int b = 0; b < 8; ++b) {
        for (int jj = 0; jj < 16; ++jj) {
            const int j = 16 * b + jj;
            if (j >= nb) break;
            double* buf = col + (j & 1) * NB;
            if (tk == jj) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int i = ti + 16 * a;
                    if (i >= j && i < nb) buf[i] = r[a][b];
                }
            }
            __syncthreads();
            double pivot = buf[j];
            if (!(pivot > 0.0)) {
                if (tid == 0 && bad == 0) bad = p.k0 + j + 1;
                pivot = nan("");
            }
            const double inv = 1.0 / pivot;
#pragma unroll
            for (int b2 = 0; b2 < 8; ++b2) {
                if (b2 < b) continue;
                const int k = tk + 16 * b2;
                if (k <= j || k >= nb) continue;
                const double ck = buf[k] * inv;
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int i = ti + 16 * a;
                    if (i >= k && i < nb) r[a][b2] = fma(-buf[i], ck, r[a][b2]);
                }
            }
            if (tk == jj) {
                const double ljj = sqrt(pivot), rinv = 1.0 / ljj;
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int i = ti + 16 * a;
                    if (i > j && i < nb) S[i * SP + j] = buf[i] * rinv;
                    else if (i == j) {
                        S[j * SP + j] = ljj;
                        invd[j] = rinv;
                        piv[j] = pivot;
                    }
                }
            }
        }
    }
    __syncthreads();

    // inverse: column jc by a lane pair; W[i][jc] kept at S[jc][i]
    {
        const int jc = tid >> 1, half = tid & 1;
        const bool colok = jc < nb;
        const double wjj = colok ? invd[jc] : 0.0;
        for (int i = 1; i < nb; ++i) {
            double acc = 0.0;
            if (colok && i > jc) {
                const double* Li = S + i * SP;
                const double* Wj = S + jc * SP;
                double acc2 = 0.0;
                int k = jc + half;
                if (k == jc) {  // first term uses the diagonal of W
                    acc = Li[k] * wjj;
                    k += 2;
                }
                for (; k + 2 < i; k += 4) {
                    acc = fma(Li[k], Wj[k], acc);
                    acc2 = fma(Li[k + 2], Wj[k + 2], acc2);
                }
                for (; k < i; k += 2) acc = fma(Li[k], Wj[k], acc);
                acc += acc2;
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if (colok && i > jc && half == 0) S[jc * SP + i] = -acc * invd[i];
            __syncwarp();
        }
    }
    __syncthreads();

    double* __restrict__ D = p.dinv + bz * p.stride_dinv;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx >> 7, j = idx & (NB - 1);
        double w = 0.0;
        if (i < nb && j < nb) {
            A[(long)i * p.lda + j] = (j <= i) ? S[i * SP + j] : 0.0;
            if (j < i) w = S[j * SP + i];
            else if (j == i) w = invd[i];
        }
        D[idx] = w;
    }
    if (tid < nb) p.logd[bz * p.stride_logd + tid] = 0.5 * log(piv[tid]);
    if (tid == 0 && bad) {
        atomicCAS(p.d_info, 0, bad);
        if (p.info_vec) atomicCAS(p.info_vec + bz, 0, bad);
    }
}

constexpr size_t DIAG_SMEM = (size_t)(NB * SP + 2 * NB + 2 * NB) * 8;

// W diagonal blocks <- dinv; everything else of W <- 0 (done by memset before)
__global__ void place_diag_kernel(const double* __restrict__ dinv, long stride_dinv, int nblk, double* __restrict__ W,
                                  long ldw, long strideW, int N) {
    const int kb = blockIdx.x, bz = blockIdx.y;
    const double* D = dinv + bz * stride_dinv + (long)kb * NB * NB;
    double* Wb = W + bz * strideW + (long)kb * NB * ldw + kb * NB;
    const int nb = min(NB, N - kb * NB);
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        const int i = idx >> 7, j = idx & (NB - 1);
        if (i < nb && j < nb) Wb[(long)i * ldw + j] = D[idx];
    }
}

}  // namespace

int launch_potrf(cudaStream_t s, const CholArgs& a) {
    if (a.N <= 0 || a.batch <= 0) return 0;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
        attr = true;
    }
    const int nblk = chol_nblk(a.N);
    const long stride_dinv = (long)nblk * NB * NB;
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB;
        const int nb = a.N - k0 < NB ? a.N - k0 : NB;
        DiagArgs d;
        d.A = a.A + (long)k0 * a.lda + k0;
        d.lda = a.lda;
        d.strideA = a.strideA;
        d.nb = nb;
        d.k0 = k0;
        d.dinv = a.dinv + (long)kb * NB * NB;
        d.stride_dinv = stride_dinv;
        d.logd = a.logd + k0;
        d.stride_logd = a.N;
        d.d_info = a.d_info;
        d.info_vec = a.info_vec;
        potrf_diag_kernel<<<a.batch, 256, DIAG_SMEM, s>>>(d);
        const int rem = a.N - k0 - nb;
        if (rem <= 0) break;
        double* panel = a.A + (long)(k0 + nb) * a.lda + k0;
        // panel <- panel * inv(L11)^T   (in place: one 128-wide column tile, so every CTA only
        // rewrites the rows it alone reads)
        GemmArgs g;
        g.transA = false; g.transB = true;
        g.M = rem; g.N = nb; g.K = nb;
        g.alpha = 1.0; g.beta = 0.0;
        g.A = panel; g.lda = a.lda; g.strideA = a.strideA;
        g.B = d.dinv; g.ldb = NB; g.strideB = stride_dinv;
        g.C = panel; g.ldc = a.lda; g.strideC = a.strideA;
        g.batch = a.batch;
        g.small_tiles = 0;
        if (launch_gemm(s, g)) return -2;
        // trailing <- trailing - panel panel^T  (lower tiles)
        GemmArgs t;
        t.transA = false; t.transB = true;
        t.M = rem; t.N = rem; t.K = nb;
        t.alpha = -1.0; t.beta = 1.0;
        t.A = panel; t.lda = a.lda; t.strideA = a.strideA;
        t.B = panel; t.ldb = a.lda; t.strideB = a.strideA;
        t.C = a.A + (long)(k0 + nb) * a.lda + (k0 + nb); t.ldc = a.lda; t.strideC = a.strideA;
        t.batch = a.batch;
        t.lower_only = 1;
        if (launch_gemm(s, t)) return -2;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int launch_trtri(cudaStream_t s, const CholArgs& a, double* W, long ldw, long strideW, double* scratch) {
    if (a.N <= 0 || a.batch <= 0) return 0;
    const int N = a.N;
    const int nblk = chol_nblk(N);
    const long stride_dinv = (long)nblk * NB * NB;
    for (int b = 0; b < a.batch; ++b)
        cudaMemsetAsync(W + b * strideW, 0, sizeof(double) * (size_t)N * ldw, s);
    place_diag_kernel<<<dim3(nblk, a.batch), 256, 0, s>>>(a.dinv, stride_dinv, nblk, W, ldw, strideW, N);
    for (long sz = NB; sz < N; sz *= 2) {
        // pairs c: first half [c*2sz, c*2sz+sz), second half [c*2sz+sz, min(c*2sz+2sz, N))
        const int nfull = (int)(N / (2 * sz));                 // pairs whose second half is full
        const int rem2 = (int)(N - nfull * 2 * sz - sz);       // rows in the ragged last second half
        for (int pass = 0; pass < 2; ++pass) {
            int npairs, s2;
            long c0;
            if (pass == 0) { npairs = nfull; s2 = (int)sz; c0 = 0; }
            else { npairs = rem2 > 0 ? 1 : 0; s2 = rem2; c0 = nfull; }
            if (npairs == 0) continue;
            const long o1 = c0 * 2 * sz;  // first-half offset of the first pair in this pass
            const long pstepL = 2 * sz * a.lda + 2 * sz, pstepW = 2 * sz * ldw + 2 * sz;
            // T = L21 * W11      (s2 x sz) = (s2 x sz) (sz x sz lower)
            GemmArgs g;
            g.transA = false; g.transB = false;
            g.M = s2; g.N = (int)sz; g.K = (int)sz;
            g.A = a.A + (o1 + sz) * a.lda + o1; g.lda = a.lda; g.strideA = pstepL; g.strideA2 = a.strideA;
            g.B = W + o1 * ldw + o1; g.ldb = ldw; g.strideB = pstepW; g.strideB2 = strideW;
            g.C = scratch + (o1 + sz) * ldw + o1; g.ldc = ldw; g.strideC = pstepW; g.strideC2 = strideW;
            g.batch = npairs; g.batch2 = a.batch;
            g.krange = KR_LO_J;
            if (launch_gemm(s, g)) return -2;
            // W21 = -W22 * T     (s2 x sz) = (s2 x s2 lower) (s2 x sz)
            GemmArgs h;
            h.transA = false; h.transB = false;
            h.M = s2; h.N = (int)sz; h.K = s2;
            h.alpha = -1.0;
            h.A = W + (o1 + sz) * ldw + (o1 + sz); h.lda = ldw; h.strideA = pstepW; h.strideA2 = strideW;
            h.B = g.C; h.ldb = ldw; h.strideB = pstepW; h.strideB2 = strideW;
            h.C = W + (o1 + sz) * ldw + o1; h.ldc = ldw; h.strideC = pstepW; h.strideC2 = strideW;
            h.batch = npairs; h.batch2 = a.batch;
            h.krange = KR_HI_I;
            if (launch_gemm(s, h)) return -2;
        }
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
