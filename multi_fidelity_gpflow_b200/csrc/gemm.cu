// gemm.cu -- batched fp64 GEMM for sm_100a on the FP64 tensor path.
//
// tcgen05.mma has no f64 kind, so the FP64 tensor path on Blackwell is the warp-level
// mma.sync.m8n8k4.f64 (SASS: DMMA.8x8x4).  This kernel feeds it from a 4-stage cp.async
// (LDGSTS, 16-byte, zero-filling) shared-memory pipeline with XOR-swizzled tiles so that the
// 8-byte fragment loads of both operand layouts are bank-conflict free:
//   * K-major operand (contiguous along k): tile [rows][16], 16-byte chunk c of row r stored at
//     chunk c ^ (r & 7); an MMA row-fragment uses rows {2g + b} of a 16-row group so the four
//     rows of a half-warp land in four different bank octets.
//   * M-major operand (contiguous along m/n): tile [16][rows], chunk c of k-row k stored at
//     chunk c ^ ((k & 3) << 1); an MMA fragment uses 8 consecutive rows.
// Warp tile 64x32 (128x128 CTA, 8 warps) or 32x32 (64x64 CTA, 4 warps): 12 (8) LDS.64 per
// 32 (16) DMMA.  Triangular operands are expressed as per-tile k ranges (GemmKRange).
#include "gemm.cuh"
#include "common.cuh"

#include <cstdint>

namespace {

constexpr int BK = 16;
constexpr int STAGES = 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- global -> shared tile copies ------------------------------------------------------------
// K-major source: element (r, k) at g[(row0 + r) * ld + k]
template <int ROWS, int NT>
__device__ __forceinline__ void load_kmajor(uint32_t sbase, const double* __restrict__ g, long ld, int row0,
                                            int nrows, int k0, int kend, int tid) {
#pragma unroll
    for (int it = 0; it < ROWS * 8 / NT; ++it) {
        const int idx = tid + it * NT;
        const int r = idx >> 3, c = idx & 7;
        const int grow = row0 + r, gk = k0 + 2 * c;
        int valid = (grow < nrows) ? (kend - gk) : 0;
        valid = valid < 0 ? 0 : (valid > 2 ? 2 : valid);
        const double* src = valid ? g + (long)grow * ld + gk : g;
        cp_async16(sbase + r * 128 + ((c ^ (r & 7)) << 4), src, valid * 8);
    }
}
// M-major source: element (r, k) at g[k * ld + row0 + r]
template <int ROWS, int NT>
__device__ __forceinline__ void load_mmajor(uint32_t sbase, const double* __restrict__ g, long ld, int row0,
                                            int nrows, int k0, int kend, int tid) {
    constexpr int CPR = ROWS / 2;  // 16-byte chunks per k-row
#pragma unroll
    for (int it = 0; it < BK * CPR / NT; ++it) {
        const int idx = tid + it * NT;
        const int k = idx / CPR, c = idx % CPR;
        const int gk = k0 + k, gi = row0 + 2 * c;
        int valid = (gk < kend) ? (nrows - gi) : 0;
        valid = valid < 0 ? 0 : (valid > 2 ? 2 : valid);
        const double* src = valid ? g + (long)gk * ld + gi : g;
        cp_async16(sbase + k * (ROWS * 8) + ((c ^ ((k & 3) << 1)) << 4), src, valid * 8);
    }
}

// tile-local row of MMA row g (0..7) of fragment f inside a warp tile starting at w0
template <bool KMAJOR>
__device__ __forceinline__ int frag_row(int w0, int f, int g) {
    return KMAJOR ? w0 + (f >> 1) * 16 + 2 * g + (f & 1) : w0 + f * 8 + g;
}
// shared address (bytes, relative to tile base) of element (row, k)
template <bool KMAJOR, int ROWS>
__device__ __forceinline__ uint32_t frag_addr(int row, int k) {
    return KMAJOR ? (uint32_t)(row * 128 + (((k >> 1) ^ (row & 7)) << 4) + ((k & 1) << 3))
                  : (uint32_t)(k * (ROWS * 8) + (((row >> 1) ^ ((k & 3) << 1)) << 4) + ((row & 1) << 3));
}

struct KernelArgs {
    int M, N, K;
    double alpha, beta;
    const double* A;
    long lda, strideA;
    const double* B;
    long ldb, strideB;
    double* C;
    long ldc, strideC;
    int krange, lower_only, vec_ok;
    int batch;
    long strideA2, strideB2, strideC2;
};

template <int BM, int BN, int WM, int WN, bool TA, bool TB>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, (BM * BN == 128 * 128 ? 1 : 2)) dgemm_kernel(KernelArgs p) {
    constexpr int NWN = BN / WN;
    constexpr int NT = (BM / WM) * NWN * 32;
    constexpr int MT = WM / 8, NTL = WN / 8;
    constexpr bool A_KM = !TA, B_KM = TB;
    constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8, STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    if (p.lower_only && j0 >= i0 + BM) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / NWN) * WM, wn0 = (warp % NWN) * WN;
    const long b1 = blockIdx.z % p.batch, b2 = blockIdx.z / p.batch;
    const double* __restrict__ A = p.A + b1 * p.strideA + b2 * p.strideA2;
    const double* __restrict__ B = p.B + b1 * p.strideB + b2 * p.strideB2;
    double* __restrict__ C = p.C + b1 * p.strideC + b2 * p.strideC2;

    int klo = 0, khi = p.K;
    switch (p.krange & 3) {
        case KR_LO_I: klo = i0; break;
        case KR_LO_J: klo = j0; break;
        case KR_LO_MAXIJ: klo = i0 > j0 ? i0 : j0; break;
        default: break;
    }
    switch (p.krange & 12) {
        case KR_HI_I: khi = min(khi, i0 + BM); break;
        case KR_HI_J: khi = min(khi, j0 + BN); break;
        case KR_HI_MINIJ: khi = min(khi, min(i0 + BM, j0 + BN)); break;
        default: break;
    }
    const int nk = (khi > klo && p.alpha != 0.0) ? (khi - klo + BK - 1) / BK : 0;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);

    auto load_stage = [&](int stage, int k0) {
        const uint32_t sa = sbase + stage * STAGE_BYTES, sb = sa + A_BYTES;
        if (A_KM) load_kmajor<BM, NT>(sa, A, p.lda, i0, p.M, k0, khi, tid);
        else load_mmajor<BM, NT>(sa, A, p.lda, i0, p.M, k0, khi, tid);
        if (B_KM) load_kmajor<BN, NT>(sb, B, p.ldb, j0, p.N, k0, khi, tid);
        else load_mmajor<BN, NT>(sb, B, p.ldb, j0, p.N, k0, khi, tid);
    };

    double acc[MT][NTL][2];
    int arow[MT], bcol[NTL];
#pragma unroll
    for (int m = 0; m < MT; ++m) arow[m] = frag_row<A_KM>(wm0, m, g);
#pragma unroll
    for (int n = 0; n < NTL; ++n) bcol[n] = frag_row<B_KM>(wn0, n, g);

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, klo + s * BK);
        cp_async_commit();
    }

#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NTL; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;

    for (int it = 0; it < nk; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = it + STAGES - 1;
        if (nxt < nk) load_stage(nxt % STAGES, klo + nxt * BK);
        cp_async_commit();
        const uint32_t sa = sbase + (it % STAGES) * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double af[MT], bf[NTL];
            const int k = kk * 4 + t;
#pragma unroll
            for (int m = 0; m < MT; ++m) af[m] = lds64(sa + frag_addr<A_KM, BM>(arow[m], k));
#pragma unroll
            for (int n = 0; n < NTL; ++n) bf[n] = lds64(sb + frag_addr<B_KM, BN>(bcol[n], k));
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < NTL; ++n) dmma(acc[m][n][0], acc[m][n][1], af[m], bf[n]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue through shared memory -----------------------------------------------------------
    // The accumulator fragments own 16-byte pieces scattered over 8 rows; reading/writing C straight from
    // them half-uses every 128-byte line and starves on outstanding-miss slots (measured: the beta*C read
    // ran at 1.6 TB/s).  Instead the C tile is moved with full-line cp.async / 128-bit stores and the
    // fragments touch only shared memory.  The pipeline buffers are free at this point.
    constexpr int CP = (BM * (BN + 2) * 8 <= STAGES * STAGE_BYTES) ? BN + 2 : BN;  // padded pitch (doubles)
    double* ct = reinterpret_cast<double*>(smem_raw);
    const double alpha = p.alpha, beta = p.beta;
    const bool full = p.vec_ok && (i0 + BM <= p.M) && (j0 + BN <= p.N);
    __syncthreads();  // every warp is done reading the last stage
    if (beta != 0.0) {
        if (full) {
#pragma unroll 4
            for (int idx = tid; idx < BM * BN / 2; idx += NT) {
                const int r = idx / (BN / 2), c = idx % (BN / 2);
                cp_async16(sbase + (r * CP + 2 * c) * 8, C + (long)(i0 + r) * p.ldc + j0 + 2 * c, 16);
            }
            cp_async_commit();
            cp_async_wait<0>();
        } else {
            for (int idx = tid; idx < BM * BN; idx += NT) {
                const int r = idx / BN, c = idx % BN;
                ct[r * CP + c] = (i0 + r < p.M && j0 + c < p.N) ? C[(long)(i0 + r) * p.ldc + j0 + c] : 0.0;
            }
        }
        __syncthreads();
    }
    auto merge2 = [&](int r, int c, double v0, double v1) {  // tile-local row r, columns c, c+1 (c even)
        double2* q = reinterpret_cast<double2*>(ct + r * CP + c);
        double2 o = make_double2(alpha * v0, alpha * v1);
        if (beta != 0.0) {
            const double2 old = *q;
            o.x = fma(beta, old.x, o.x);
            o.y = fma(beta, old.y, o.y);
        }
        *q = o;
    };
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        const int r = arow[m];
        if (B_KM) {
            // fragment pair (2q, 2q+1) holds columns 16q + 4t + {0,1,2,3}
#pragma unroll
            for (int q = 0; q < NTL / 2; ++q) {
                const int c = wn0 + q * 16 + 4 * t;
                merge2(r, c, acc[m][2 * q][0], acc[m][2 * q + 1][0]);
                merge2(r, c + 2, acc[m][2 * q][1], acc[m][2 * q + 1][1]);
            }
        } else {
#pragma unroll
            for (int n = 0; n < NTL; ++n) merge2(r, wn0 + n * 8 + 2 * t, acc[m][n][0], acc[m][n][1]);
        }
    }
    __syncthreads();
    if (full) {
#pragma unroll 4
        for (int idx = tid; idx < BM * BN / 2; idx += NT) {
            const int r = idx / (BN / 2), c = idx % (BN / 2);
            *reinterpret_cast<double2*>(C + (long)(i0 + r) * p.ldc + j0 + 2 * c) =
                *reinterpret_cast<const double2*>(ct + r * CP + 2 * c);
        }
    } else {
        for (int idx = tid; idx < BM * BN; idx += NT) {
            const int r = idx / BN, c = idx % BN;
            if (i0 + r < p.M && j0 + c < p.N) C[(long)(i0 + r) * p.ldc + j0 + c] = ct[r * CP + c];
        }
    }
}

template <int BM, int BN, int WM, int WN, bool TA, bool TB>
int launch_cfg(cudaStream_t s, const GemmArgs& a, const KernelArgs& ka) {
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int SMEM = STAGES * (BM + BN) * BK * 8;
    static SmemOptIn optin;
    if (!optin.ensure(dgemm_kernel<BM, BN, WM, WN, TA, TB>, SMEM)) return -2;
    dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, a.batch * a.batch2);
    dgemm_kernel<BM, BN, WM, WN, TA, TB><<<grid, NT, SMEM, s>>>(ka);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

template <bool TA, bool TB>
int launch_t(cudaStream_t s, const GemmArgs& a, const KernelArgs& ka, int cfg) {
    if (cfg == 1) return launch_cfg<64, 64, 32, 32, TA, TB>(s, a, ka);
    if (cfg == 2) return launch_cfg<128, 64, 64, 32, TA, TB>(s, a, ka);
    return launch_cfg<128, 128, 64, 32, TA, TB>(s, a, ka);
}

}  // namespace

int launch_gemm(cudaStream_t s, const GemmArgs& a) {
    if (a.M <= 0 || a.N <= 0 || a.batch <= 0 || a.batch2 <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<size_t>(p) & 15) == 0; };
    if (!al16(a.A) || !al16(a.B) || (a.lda & 1) || (a.ldb & 1) || (a.strideA & 1) || (a.strideB & 1) || (a.strideA2 & 1) || (a.strideB2 & 1)) return -1;
    KernelArgs ka;
    ka.M = a.M; ka.N = a.N; ka.K = a.K;
    ka.alpha = a.alpha; ka.beta = a.beta;
    ka.A = a.A; ka.lda = a.lda; ka.strideA = a.strideA;
    ka.B = a.B; ka.ldb = a.ldb; ka.strideB = a.strideB;
    ka.C = a.C; ka.ldc = a.ldc; ka.strideC = a.strideC;
    ka.krange = a.krange; ka.lower_only = a.lower_only;
    ka.batch = a.batch; ka.strideA2 = a.strideA2; ka.strideB2 = a.strideB2; ka.strideC2 = a.strideC2;
    ka.vec_ok = al16(a.C) && !(a.ldc & 1) && !(a.strideC & 1) && !(a.strideC2 & 1);
    // tile config: 0 = 128x128 (1 CTA/SM; the only one safe for the in-place panel solve), 1 = 64x64,
    // 2 = 128x64 (2 CTAs/SM: one CTA's prologue/epilogue hides behind the other's main loop) -- the default.
    int cfg;
    if (a.small_tiles >= 0) cfg = a.small_tiles;
    else {
        const long ctas = (long)((a.M + 127) / 128) * ((a.N + 63) / 64) * a.batch * a.batch2;
        cfg = ctas < 2 * 148 ? 1 : 2;
    }
    if (!a.transA && !a.transB) return launch_t<false, false>(s, a, ka, cfg);
    if (!a.transA && a.transB) return launch_t<false, true>(s, a, ka, cfg);
    if (a.transA && !a.transB) return launch_t<true, false>(s, a, ka, cfg);
    return launch_t<true, true>(s, a, ka, cfg);
}
