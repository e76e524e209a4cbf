// gpr_small.cu -- K6: "one GP per k-bin" batched small-matrix NLML + analytic gradient.
//
// One CTA (128 threads) per problem, everything resident in shared memory / registers:
//   assemble K (fused MF kernel)  ->  register-resident right-looking Cholesky  ->  L^-1
//   ->  a = L^-1 y, alpha = L^-T a  ->  K^-1 = L^-T L^-1 (never stored)  ->
//   G = alpha alpha^T - K^-1  contracted with dK/dtheta recomputed on the fly.
// Replaces, per bin, GPR.log_marginal_likelihood + tape.gradient (reference
// mfgpflow/linear.py:206-207) for the many-single-output-GP layout the reference describes in
// gpemulator_singlebin.py:1-14.  N <= 64, P = 1 per problem.
//
// Ownership map for all N x N triangular work: thread (ti, tk) = (tid / 16, tid % 16) owns the
// elements (i, k) = (ti + 8a, tk + 16b), a < 8, b < 4, k <= i -- a 2-D cyclic layout, so the
// shrinking trailing matrix of the Cholesky stays balanced across the CTA and each step needs
// one published column (double-buffered) and ONE __syncthreads.
#include "gpr_small.cuh"

#include <cmath>

namespace {

constexpr int NTH = 128;
constexpr double LOG2PI = 1.8378770664093454835606594728112;

__device__ inline double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Smem {
    double *Ks, *Ws, *cL, *cD, *sv, *hv, *yv, *av, *al, *invd, *piv, *col, *th, *red;
    int* hidx;
    int* misc;  // [0] bad pivot (1-based), [1] number of HF points
};

__host__ __device__ inline size_t smem_doubles(int N, int d, int NP) {
    return (size_t)2 * N * NP + 2 * d * N + 7 * 64 + 2 * 64 + (4 * d + 4) + 4 * (2 * d + 4) + 64 /*hidx ints*/ + 2;
}

__global__ void __launch_bounds__(NTH) gpr_small_kernel(SmallArgs p) {
    extern __shared__ __align__(16) double smem[];
    const int N = p.N, d = p.d, NP = p.NP;
    const int prob = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ti = tid >> 4, tk = tid & 15;
    Smem s;
    s.Ks = smem;
    s.Ws = s.Ks + N * NP;
    s.cL = s.Ws + N * NP;
    s.cD = s.cL + d * N;
    s.sv = s.cD + d * N;
    s.hv = s.sv + 64;
    s.yv = s.hv + 64;
    s.av = s.yv + 64;
    s.al = s.av + 64;
    s.invd = s.al + 64;
    s.piv = s.invd + 64;
    s.col = s.piv + 64;  // [2][64]
    s.th = s.col + 128;  // theta [2d+3], inverse length-scales [2d]
    s.red = s.th + 4 * d + 4;
    s.hidx = reinterpret_cast<int*>(s.red + 4 * (2 * d + 4));
    s.misc = s.hidx + 64;

    const double* __restrict__ theta = p.theta + (long)prob * (2 * d + 3);
    if (tid < 2 * d + 3) s.th[tid] = theta[tid];
    if (tid == 0) {
        s.misc[0] = 0;
        s.misc[1] = 0;
    }
    __syncthreads();
    if (tid < 2 * d) s.th[2 * d + 3 + tid] = 1.0 / (tid < d ? s.th[1 + tid] : s.th[2 + tid]);
    __syncthreads();
    const double rho = s.th[0], vL = s.th[1 + d], vD = s.th[2 + 2 * d];
    const double noise = p.noise[prob];
    if (tid < N) {
        const double fid = p.X[(long)tid * (d + 1) + d];
        double sf = 0.0, hf = 0.0;
        bool live = false;
        if (fid == 0.0) { sf = 1.0; live = true; }
        else if (fid == 1.0) { sf = rho; hf = 1.0; live = true; }
        for (int q = 0; q < d; ++q) {
            const double x = live ? p.X[(long)tid * (d + 1) + q] : 0.0;
            s.cL[q * N + tid] = x * s.th[2 * d + 3 + q];
            s.cD[q * N + tid] = x * s.th[3 * d + 3 + q];
        }
        s.sv[tid] = sf;
        s.hv[tid] = hf;
        s.yv[tid] = p.Y[(long)tid * p.ldy + prob % p.ycols];
    }
    __syncthreads();
    if (tid == 0) {  // ordered list of HF points (used by the discrepancy-kernel gradient pass)
        int n = 0;
        for (int i = 0; i < N; ++i)
            if (s.hv[i] != 0.0) s.hidx[n++] = i;
        s.misc[1] = n;
    }

    // ---- phase A: assemble the owned elements of K (registers) and K^L (Ws upper) ------------
    double r[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = ti + 8 * a, k = tk + 16 * b;
            double v = 0.0;
            if (k <= i && i < N) {
                double rl = 0.0;
                for (int q = 0; q < d; ++q) {
                    const double df = s.cL[q * N + i] - s.cL[q * N + k];
                    rl = fma(df, df, rl);
                }
                const double kl = s.sv[i] * s.sv[k] * vL * exp(-0.5 * rl);
                v = kl;
                if (s.hv[i] != 0.0 && s.hv[k] != 0.0) {
                    double rd = 0.0;
                    for (int q = 0; q < d; ++q) {
                        const double df = s.cD[q * N + i] - s.cD[q * N + k];
                        rd = fma(df, df, rd);
                    }
                    v = fma(vD, exp(-0.5 * rd), v);
                }
                if (i == k) v += noise;
                else s.Ws[k * NP + i] = kl;
            }
            r[a][b] = v;
        }

    // ---- phase B: right-looking Cholesky, trailing matrix in registers -----------------------
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        for (int jj = 0; jj < 16; ++jj) {
            const int j = 16 * b + jj;
            if (j >= N) break;
            double* buf = s.col + (j & 1) * 64;
            if (tk == jj) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int i = ti + 8 * a;
                    if (i >= j && i < N) buf[i] = r[a][b];
                }
            }
            __syncthreads();
            double pivot = buf[j];
            if (!(pivot > 0.0)) {
                if (tid == 0 && s.misc[0] == 0) s.misc[0] = j + 1;
                pivot = nan("");
            }
            const double inv = 1.0 / pivot;
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) {
                if (b2 < b) continue;
                const int k = tk + 16 * b2;
                if (k <= j || k >= N) continue;
                const double ck = buf[k] * inv;
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int i = ti + 8 * a;
                    if (i >= k && i < N) r[a][b2] = fma(-buf[i], ck, r[a][b2]);
                }
            }
            if (tk == jj) {  // owners of column j write the final L column
                const double ljj = sqrt(pivot), rinv = 1.0 / ljj;
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int i = ti + 8 * a;
                    if (i > j && i < N) s.Ks[i * NP + j] = buf[i] * rinv;
                    else if (i == j) {
                        s.Ks[j * NP + j] = ljj;
                        s.invd[j] = rinv;
                        s.piv[j] = pivot;
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- phase C: W = L^-1, column jc by a lane pair; W[i][jc] stored at Ks[jc][i] (upper) ----
    {
        const int jc = tid >> 1, half = tid & 1;
        const bool colok = jc < N;
        const double wjj = colok ? s.invd[jc] : 0.0;
        for (int i = 1; i < N; ++i) {
            double acc = 0.0;
            if (colok && i > jc) {
                const double* Li = s.Ks + i * NP;
                const double* Wj = s.Ks + jc * NP;
                for (int k = jc + half; k < i; k += 2) acc = fma(Li[k], (k == jc) ? wjj : Wj[k], acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if (colok && i > jc && half == 0) s.Ks[jc * NP + i] = -acc * s.invd[i];
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- phase D: a = W y, alpha = W^T a, value ----------------------------------------------
    if (tid < N) {
        double acc = s.invd[tid] * s.yv[tid];
        for (int k = 0; k < tid; ++k) acc = fma(s.Ks[k * NP + tid], s.yv[k], acc);
        s.av[tid] = acc;
    }
    __syncthreads();
    if (tid < N) {
        double acc = s.invd[tid] * s.av[tid];
        for (int i = tid + 1; i < N; ++i) acc = fma(s.Ks[tid * NP + i], s.av[i], acc);
        s.al[tid] = acc;
    }
    if (warp == 3) {  // nlml = 0.5 |a|^2 + sum log L_ii + N/2 log 2pi
        double q = 0.0, ld = 0.0;
        for (int i = lane; i < N; i += 32) {
            q = fma(s.av[i], s.av[i], q);
            ld += log(s.piv[i]);
        }
        q = warp_sum(q);
        ld = warp_sum(ld);
        if (lane == 0) {
            const int bad = s.misc[0];
            p.nlml[prob] = 0.5 * q + 0.5 * ld + 0.5 * N * LOG2PI;
            if (p.info) p.info[prob] = bad;
            if (bad) atomicCAS(p.d_info, 0, bad);
        }
    }
    if (!p.grad) return;
    __syncthreads();

    // ---- phase E: K^-1 on the owned elements, G = alpha alpha^T - K^-1, contraction ----------
    const int nq = 2 * d + 4;
    double acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int k = ti + 1; k < N; ++k) {
        double rj[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = tk + 16 * b;
            rj[b] = (j < k) ? s.Ks[j * NP + k] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int i = ti + 8 * a;
            if (i < k) {
                const double ri = s.Ks[i * NP + k];
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(ri, rj[b], acc[a][b]);
            }
        }
    }
    double s_vL = 0.0, s_rho = 0.0, s_dg = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = ti + 8 * a, j = tk + 16 * b;
            double tl = 0.0;
            if (j <= i && i < N) {
                const double wij = (i == j) ? s.invd[i] : s.Ks[j * NP + i];
                const double kinv = fma(s.invd[i], wij, acc[a][b]);
                const double g = s.al[i] * s.al[j] - kinv;
                s.Ws[i * NP + j] = g;  // lower triangle of Ws <- G (read by the HF x HF pass)
                if (i == j) {
                    s_dg += g;
                    tl = g * s.sv[i] * s.sv[i] * vL;
                } else {
                    tl = 2.0 * g * s.Ws[j * NP + i];
                }
                s_vL += tl;
                s_rho += tl * (s.hv[i] + s.hv[j]);
            }
            acc[a][b] = tl;  // now T^L = w G K^L
        }
    double* myred = s.red + warp * nq;
    {
        const double a0 = warp_sum(s_rho), a1 = warp_sum(s_vL), a2 = warp_sum(s_dg);
        if (lane == 0) {
            myred[0] = a0;
            myred[1 + d] = a1;
            myred[3 + 2 * d] = a2;
        }
    }
    for (int q = 0; q < d; ++q) {
        double sL = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = tk + 16 * b;
            const double xj = (j < N) ? s.cL[q * N + j] : 0.0;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int i = ti + 8 * a;
                if (j <= i && i < N) {
                    const double df = s.cL[q * N + i] - xj;
                    sL = fma(acc[a][b] * df, df, sL);
                }
            }
        }
        sL = warp_sum(sL);
        if (lane == 0) myred[1 + q] = sL;
    }
    __syncthreads();  // G complete in Ws lower
    {
        // discrepancy kernel: only HF x HF pairs contribute (h_i h_j = 1)
        const int nH = s.misc[1];
        const int npairs = nH * (nH + 1) / 2;
        double s_vD = 0.0;
        // per-thread list walk; the per-dimension sums are accumulated in a second walk below
        for (int t = tid; t < npairs; t += NTH) {
            int pi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
            while ((pi + 1) * (pi + 2) / 2 <= t) ++pi;
            while (pi * (pi + 1) / 2 > t) --pi;
            const int pj = t - pi * (pi + 1) / 2;
            const int i = s.hidx[pi], j = s.hidx[pj];  // i >= j
            double rd = 0.0;
            for (int q = 0; q < d; ++q) {
                const double df = s.cD[q * N + i] - s.cD[q * N + j];
                rd = fma(df, df, rd);
            }
            const double tdv = ((i == j) ? 1.0 : 2.0) * s.Ws[i * NP + j] * vD * exp(-0.5 * rd);
            s_vD += tdv;
        }
        s_vD = warp_sum(s_vD);
        if (lane == 0) myred[2 + 2 * d] = s_vD;
        for (int q = 0; q < d; ++q) {
            double sD = 0.0;
            for (int t = tid; t < npairs; t += NTH) {
                int pi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
                while ((pi + 1) * (pi + 2) / 2 <= t) ++pi;
                while (pi * (pi + 1) / 2 > t) --pi;
                const int pj = t - pi * (pi + 1) / 2;
                if (pi == pj) continue;
                const int i = s.hidx[pi], j = s.hidx[pj];
                double rd = 0.0;
                for (int qq = 0; qq < d; ++qq) {
                    const double df = s.cD[qq * N + i] - s.cD[qq * N + j];
                    rd = fma(df, df, rd);
                }
                const double df = s.cD[q * N + i] - s.cD[q * N + j];
                sD = fma(2.0 * s.Ws[i * NP + j] * vD * exp(-0.5 * rd) * df, df, sD);
            }
            sD = warp_sum(sD);
            if (lane == 0) myred[2 + d + q] = sD;
        }
    }
    __syncthreads();
    if (tid < nq) {
        double tot = s.red[tid] + s.red[nq + tid] + s.red[2 * nq + tid] + s.red[3 * nq + tid];
        double f = 1.0;
        if (tid == 0) f = 1.0 / rho;
        else if (tid <= d) f = s.th[2 * d + 3 + (tid - 1)];
        else if (tid == d + 1) f = 1.0 / vL;
        else if (tid <= 2 * d + 1) f = s.th[3 * d + 3 + (tid - d - 2)];
        else if (tid == 2 * d + 2) f = 1.0 / vD;
        p.grad[(long)prob * nq + tid] = -0.5 * f * tot;  // d(nlml) = -1/2 sum G dK
    }
}

}  // namespace

int launch_gpr_small(cudaStream_t st, const SmallArgs& a0) {
    SmallArgs a = a0;
    if (a.N < 1 || a.N > 64 || a.d < 1 || a.d > MFGP_SMALL_MAX_D) return -1;
    if (a.B <= 0) return 0;
    a.NP = a.N | 1;
    const size_t bytes = smem_doubles(a.N, a.d, a.NP) * 8;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(gpr_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(smem_doubles(64, MFGP_SMALL_MAX_D, 65) * 8));
        attr = true;
    }
    gpr_small_kernel<<<a.B, NTH, bytes, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
