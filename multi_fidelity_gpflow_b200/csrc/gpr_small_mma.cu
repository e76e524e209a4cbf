// gpr_small_mma.cu -- K6 v2: batched small-matrix NLML + analytic gradient, ONE WARP PER PROBLEM,
// every O(N^3) step on the FP64 tensor path (mma.sync m8n8k4 f64 -> DMMA.8x8x4).
//
// The N x N problem (N <= 64, e.g. the 53-point HBS k-bin GPs) is padded to NT x NT tiles of 8x8.
// Lower-triangular tiles live in a per-warp shared-memory array of XOR-swizzled 512-byte tiles,
// laid out so that the three access patterns a DMMA fragment needs are bank-conflict free:
//   C-fragment store (row g, cols 2t,2t+1 -> one STS.128), K-major fragment (row g, col t+4s) and
//   M-major fragment (row t+4s, col g):   byte(r, c) = r*64 + (((c>>1) ^ (r&2)) << 4) + ((c&1) << 3).
// A warp owns its problem, so only __syncwarp orders its smem traffic; the few __syncthreads keep the
// 4 warps of a CTA (identical instruction streams) in step so they share instruction-cache lines.
//
//   A+B  left-looking blocked Cholesky; block column kb is ASSEMBLED on the fly into DMMA accumulators
//        (fused MF kernel, expanded-square distance + exp), updated with sum_k L_ik L_kbk^T (DMMA), the
//        8x8 diagonal tile is factored redundantly in registers (no shuffles) and inverted, and the
//        panel below is solved as A_ik inv(L_kk)^T (DMMA).  The diagonal slot keeps inv(L_kk).
//   C    W = L^-1 in place, W_ij = -W_ii sum_k L_ik W_kj (DMMA)
//   D    a = W y, alpha = W^T a as one-column DMMA products; nlml = 1/2|a|^2 + sum log L_ii + N/2 log 2pi
//   E    K^-1 = W^T W accumulated in registers (DMMA), G = alpha alpha^T - K^-1, contraction with
//        dK/dtheta recomputed on the fly (K^L re-evaluated; only HF x HF pairs touch the delta kernel).
// Replaces per bin GPR.log_marginal_likelihood + tape.gradient (reference mfgpflow/linear.py:206-207).
#include <cmath>
#include <cstdint>

#include "gpr_small.cuh"
#include "mathx.cuh"

namespace {

constexpr int WPC = 4;  // warps (= problems) per CTA (2 CTAs per SM; 3x3 and 5x2 were measured slower)
constexpr double LOG2PI = 1.8378770664093454835606594728112;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ constexpr int slot(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j
__device__ __forceinline__ int tile_off(int r, int c) {  // element offset (doubles) inside a 64-double tile
    return r * 8 + ((((c >> 1) ^ (r & 2))) << 1) + (c & 1);
}

struct Lane {
    int g, t;
    int cst;     // C-fragment store offset (doubles): row g, cols 2t,2t+1
    int km[2];   // K-major fragment offsets: (row g, col t + 4s)
    int mm[2];   // M-major fragment offsets: (row t + 4s, col g)
};

__device__ __forceinline__ void st_c(double* tile, const Lane& L, double c0, double c1) {
    *reinterpret_cast<double2*>(tile + L.cst) = make_double2(c0, c1);
}

// per-warp shared memory carve-up (doubles)
template <int NT>
struct WarpMem {
    static constexpr int NP = 8 * NT;
    static constexpr int NTRI = NT * (NT + 1) / 2;
    double* tiles;  // [NTRI][64]
    double *xL, *xD;  // [d][NP]
    double *hL, *hD, *sv, *hv, *yv, *av, *al;  // [NP]
    double* th;   // theta[2d+3], inverse length-scales [2d]
    double* red;  // [2d+4]
    int* hidx;    // [NP]
    __host__ __device__ static size_t doubles(int d) {
        return (size_t)NTRI * 64 + 2 * (size_t)d * NP + 7 * NP + (4 * d + 4) + (2 * d + 4) + NP / 2 + 2;
    }
    __device__ WarpMem(double* base, int d) {
        tiles = base;
        xL = tiles + NTRI * 64;
        xD = xL + d * NP;
        hL = xD + d * NP;
        hD = hL + NP;
        sv = hD + NP;
        hv = sv + NP;
        yv = hv + NP;
        av = yv + NP;
        al = av + NP;
        th = al + NP;
        red = th + 4 * d + 4;
        hidx = reinterpret_cast<int*>(red + 2 * d + 4);
    }
};

template <int NT>
__global__ void __launch_bounds__(WPC * 32) gpr_small_mma_kernel(SmallArgs p, size_t warp_doubles) {
    constexpr int NP = 8 * NT;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prob_raw = blockIdx.x * WPC + warp;
    const bool active = prob_raw < p.B;
    const int prob = active ? prob_raw : p.B - 1;  // idle warps shadow the last problem (keeps __syncthreads legal)
    const int N = p.N, d = p.d;
    WarpMem<NT> m(smem + (size_t)warp * warp_doubles, d);
    Lane L;
    L.g = lane >> 2;
    L.t = lane & 3;
    L.cst = tile_off(L.g, 2 * L.t);
    L.km[0] = tile_off(L.g, L.t);
    L.km[1] = tile_off(L.g, L.t + 4);
    L.mm[0] = tile_off(L.t, L.g);
    L.mm[1] = tile_off(L.t + 4, L.g);
    const int g = L.g, t = L.t;

    // ---- setup: theta, scaled coordinates, fidelity factors, y ---------------------------------
    const double* __restrict__ theta = p.theta + (size_t)prob * (2 * d + 3);
    for (int q = lane; q < 2 * d + 3; q += 32) m.th[q] = theta[q];
    __syncwarp();
    for (int q = lane; q < 2 * d; q += 32) m.th[2 * d + 3 + q] = 1.0 / (q < d ? m.th[1 + q] : m.th[2 + q]);
    __syncwarp();
    const double rho = m.th[0], vL = m.th[1 + d], vD = m.th[2 + 2 * d];
    const double noise = p.noise[prob];
    unsigned hmask = 0;  // bit i: tile row i contains an HF point
    for (int r = lane; r < NP; r += 32) {
        double sf = 0.0, hf = 0.0;
        bool live = false;
        if (r < N) {
            const double fid = p.X[(size_t)r * (d + 1) + d];
            if (fid == 0.0) { sf = 1.0; live = true; }
            else if (fid == 1.0) { sf = rho; hf = 1.0; live = true; }
        }
        double nL = 0.0, nD = 0.0;
        for (int q = 0; q < d; ++q) {
            const double x = live ? p.X[(size_t)r * (d + 1) + q] : 0.0;
            const double xl = x * m.th[2 * d + 3 + q], xd = x * m.th[3 * d + 3 + q];
            m.xL[q * NP + r] = xl;
            m.xD[q * NP + r] = xd;
            nL = fma(xl, xl, nL);
            nD = fma(xd, xd, nD);
        }
        m.hL[r] = -0.5 * nL;
        m.hD[r] = -0.5 * nD;
        m.sv[r] = sf;
        m.hv[r] = hf;
        m.yv[r] = (r < N) ? p.Y[(size_t)r * p.ldy + prob % p.ycols] : 0.0;
        if (hf != 0.0) hmask |= 1u << (r >> 3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hmask |= __shfl_xor_sync(0xffffffffu, hmask, o);
    if (lane == 0) {  // ordered list of HF points for the discrepancy-kernel gradient pass
        int n = 0;
        for (int i = 0; i < N; ++i)
            if (p.X[(size_t)i * (d + 1) + d] == 1.0) m.hidx[n++] = i;
        m.red[0] = (double)n;  // stash the count (read back after the barrier below)
    }
    __syncwarp();
    const int nH = (int)m.red[0];
    __syncwarp();

    int bad = 0;           // first non-positive pivot (1-based), identical on all lanes
    double logdet2 = 0.0;  // sum log(pivot) = 2 sum log L_ii

    // =============================== phases A + B ================================================
#pragma unroll
    for (int kb = 0; kb < NT; ++kb) {
        __syncthreads();  // instruction-stream re-alignment (I-cache sharing)
        double acc[NT][2];
        // ---- assemble block column kb: tiles (i, kb), i >= kb ------------------------------------
        {
            double e[NT][2];
#pragma unroll
            for (int i = kb; i < NT; ++i) e[i][0] = e[i][1] = 0.0;
            for (int q = 0; q < d; ++q) {
                const double2 xc = *reinterpret_cast<const double2*>(m.xL + q * NP + 8 * kb + 2 * t);
#pragma unroll
                for (int i = kb; i < NT; ++i) {
                    const double xr = m.xL[q * NP + 8 * i + g];
                    e[i][0] = fma(xr, xc.x, e[i][0]);
                    e[i][1] = fma(xr, xc.y, e[i][1]);
                }
            }
            const int c0 = 8 * kb + 2 * t;
            const double2 hc = *reinterpret_cast<const double2*>(m.hL + c0);
            const double2 sc = *reinterpret_cast<const double2*>(m.sv + c0);
#pragma unroll
            for (int i = kb; i < NT; ++i) {
                const int r = 8 * i + g;
                const double hr = m.hL[r], sr = m.sv[r];
                acc[i][0] = (sr * sc.x) * vL * fexp(e[i][0] + (hr + hc.x));
                acc[i][1] = (sr * sc.y) * vL * fexp(e[i][1] + (hr + hc.y));
            }
            if ((hmask >> kb) & 1u) {  // this block column has HF points: add the discrepancy GP on HF x HF
#pragma unroll
                for (int i = kb; i < NT; ++i) e[i][0] = e[i][1] = 0.0;
                for (int q = 0; q < d; ++q) {
                    const double2 xc = *reinterpret_cast<const double2*>(m.xD + q * NP + 8 * kb + 2 * t);
#pragma unroll
                    for (int i = kb; i < NT; ++i) {
                        const double xr = m.xD[q * NP + 8 * i + g];
                        e[i][0] = fma(xr, xc.x, e[i][0]);
                        e[i][1] = fma(xr, xc.y, e[i][1]);
                    }
                }
                const double2 hdc = *reinterpret_cast<const double2*>(m.hD + c0);
                const double2 hvc = *reinterpret_cast<const double2*>(m.hv + c0);
#pragma unroll
                for (int i = kb; i < NT; ++i) {
                    if (!((hmask >> i) & 1u)) continue;
                    const int r = 8 * i + g;
                    const double hr = m.hD[r], hvr = m.hv[r];
                    if (hvr * hvc.x != 0.0) acc[i][0] += vD * fexp(e[i][0] + (hr + hdc.x));
                    if (hvr * hvc.y != 0.0) acc[i][1] += vD * fexp(e[i][1] + (hr + hdc.y));
                }
            }
            // diagonal: + noise (real rows) or identity (padding rows keep the factorisation well posed)
            {
                const int r = 8 * kb + g;
                if (r == c0) acc[kb][0] += (r < N) ? noise : 1.0;
                if (r == c0 + 1) acc[kb][1] += (r < N) ? noise : 1.0;
            }
        }
        // ---- left-looking update: acc[i] -= sum_{k<kb} L_ik L_kbk^T ------------------------------
#pragma unroll
        for (int k = 0; k < kb; ++k) {
            const double* tb = m.tiles + slot(kb, k) * 64;
            const double b0 = tb[L.km[0]], b1 = tb[L.km[1]];
#pragma unroll
            for (int i = kb; i < NT; ++i) {
                const double* ta = m.tiles + slot(i, k) * 64;
                const double a0 = (i == kb) ? b0 : ta[L.km[0]];
                const double a1 = (i == kb) ? b1 : ta[L.km[1]];
                dmma(acc[i][0], acc[i][1], -a0, b0);
                dmma(acc[i][0], acc[i][1], -a1, b1);
            }
        }
        // ---- diagonal tile: publish, factor redundantly in registers, invert ---------------------
        double* td = m.tiles + slot(kb, kb) * 64;
        st_c(td, L, acc[kb][0], acc[kb][1]);
#pragma unroll
        for (int i = kb + 1; i < NT; ++i) st_c(m.tiles + slot(i, kb) * 64, L, acc[i][0], acc[i][1]);
        __syncwarp();
        double af0[NT], af1[NT];  // K-major fragments of the raw panel tiles (loaded before they are overwritten)
#pragma unroll
        for (int i = kb + 1; i < NT; ++i) {
            const double* ta = m.tiles + slot(i, kb) * 64;
            af0[i] = ta[L.km[0]];
            af1[i] = ta[L.km[1]];
        }
        {
            double a[8][8], rinv[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) a[r][c] = td[tile_off(r, c)];
            double prod = 1.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                double piv = a[j][j];
                if (!(piv > 0.0)) {
                    if (!bad) bad = 8 * kb + j + 1;
                    piv = nan("");
                }
                prod *= piv;
                const double ri = rsqrt(piv);
                rinv[j] = ri;
#pragma unroll
                for (int i = j + 1; i < 8; ++i) a[i][j] *= ri;
#pragma unroll
                for (int k = j + 1; k < 8; ++k)
#pragma unroll
                    for (int i = k; i < 8; ++i) a[i][k] = fma(-a[i][j], a[k][j], a[i][k]);
            }
            logdet2 += log(prod);
            // column c = lane & 7 of inv(L_kk) by forward substitution (uniform control flow)
            const int c = lane & 7;
            double w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) s = fma(-a[i][k], w[k], s);
                w[i] = s * rinv[i];
            }
            __syncwarp();  // every lane has read the raw diagonal tile
            if (lane < 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) td[tile_off(i, c)] = w[i];
            }
        }
        __syncwarp();
        // ---- panel: L_ik = A_ik inv(L_kk)^T -------------------------------------------------------
        if (kb + 1 < NT) {
            const double wb0 = td[L.km[0]], wb1 = td[L.km[1]];
#pragma unroll
            for (int i = kb + 1; i < NT; ++i) {
                double c0 = 0.0, c1 = 0.0;
                dmma(c0, c1, af0[i], wb0);
                dmma(c0, c1, af1[i], wb1);
                st_c(m.tiles + slot(i, kb) * 64, L, c0, c1);
            }
        }
        __syncwarp();
    }

    // =============================== phase C: W = L^-1 in place ===================================
#pragma unroll
    for (int j = 0; j + 1 < NT; ++j) {
        __syncthreads();
#pragma unroll
        for (int i = j + 1; i < NT; ++i) {
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int k = j; k < i; ++k) {
                const double* ta = m.tiles + slot(i, k) * 64;  // L_ik (K-major: row g, col t+4s)
                const double* tb = m.tiles + slot(k, j) * 64;  // W_kj (M-major: row t+4s, col g)
                dmma(c0, c1, ta[L.km[0]], tb[L.mm[0]]);
                dmma(c0, c1, ta[L.km[1]], tb[L.mm[1]]);
            }
            double* tij = m.tiles + slot(i, j) * 64;
            __syncwarp();
            st_c(tij, L, c0, c1);  // T = sum_k L_ik W_kj  (L_ij itself is dead from here on)
            __syncwarp();
            const double* tw = m.tiles + slot(i, i) * 64;  // W_ii
            double w0 = 0.0, w1 = 0.0;
            dmma(w0, w1, -tw[L.km[0]], tij[L.mm[0]]);
            dmma(w0, w1, -tw[L.km[1]], tij[L.mm[1]]);
            __syncwarp();
            st_c(tij, L, w0, w1);
            __syncwarp();
        }
    }

    // =============================== phase D: a = W y, alpha = W^T a, value ========================
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const double* tw = m.tiles + slot(i, j) * 64;
            const double y0 = (g == 0) ? m.yv[8 * j + t] : 0.0, y1 = (g == 0) ? m.yv[8 * j + t + 4] : 0.0;
            dmma(c0, c1, tw[L.km[0]], y0);
            dmma(c0, c1, tw[L.km[1]], y1);
        }
        if (t == 0) m.av[8 * i + g] = c0;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int i = j; i < NT; ++i) {
            const double* tw = m.tiles + slot(i, j) * 64;  // A[m][k] = W_ij[k][m]
            const double a0 = (g == 0) ? m.av[8 * i + t] : 0.0, a1 = (g == 0) ? m.av[8 * i + t + 4] : 0.0;
            dmma(c0, c1, tw[L.mm[0]], a0);
            dmma(c0, c1, tw[L.mm[1]], a1);
        }
        if (t == 0) m.al[8 * j + g] = c0;
    }
    {
        double q = 0.0;
        for (int r = lane; r < NP; r += 32) q = fma(m.av[r], m.av[r], q);
        q = warp_sum(q);
        if (lane == 0 && active) {
            p.nlml[prob] = 0.5 * q + 0.5 * logdet2 + 0.5 * N * LOG2PI;
            if (p.info) p.info[prob] = bad;
            if (bad) atomicCAS(p.d_info, 0, bad);
        }
    }
    if (!p.grad) return;
    __syncwarp();

    // =============================== phase E: K^-1, G, contraction ===================================
    constexpr int NTRI = NT * (NT + 1) / 2;
    double kacc[NTRI][2];
#pragma unroll
    for (int s = 0; s < NTRI; ++s) kacc[s][0] = kacc[s][1] = 0.0;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        __syncthreads();
        double f0[NT], f1[NT];
#pragma unroll
        for (int i = 0; i <= k; ++i) {
            const double* tw = m.tiles + slot(k, i) * 64;  // W_ki, M-major: serves as A (W_ki^T) and as B (W_kj)
            f0[i] = tw[L.mm[0]];
            f1[i] = tw[L.mm[1]];
        }
#pragma unroll
        for (int i = 0; i <= k; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                dmma(kacc[slot(i, j)][0], kacc[slot(i, j)][1], f0[i], f0[j]);
                dmma(kacc[slot(i, j)][0], kacc[slot(i, j)][1], f1[i], f1[j]);
            }
    }
    __syncthreads();  // W tiles are dead: the slots are reused for G
    const int nq = 2 * d + 4;
    double s_vL = 0.0, s_rho = 0.0, s_dg = 0.0;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        __syncthreads();
        // K^L of block column j recomputed exactly as in the assembly (expanded-square form)
        double e[NT][2];
#pragma unroll
        for (int i = j; i < NT; ++i) e[i][0] = e[i][1] = 0.0;
        for (int q = 0; q < d; ++q) {
            const double2 xc = *reinterpret_cast<const double2*>(m.xL + q * NP + 8 * j + 2 * t);
#pragma unroll
            for (int i = j; i < NT; ++i) {
                const double xr = m.xL[q * NP + 8 * i + g];
                e[i][0] = fma(xr, xc.x, e[i][0]);
                e[i][1] = fma(xr, xc.y, e[i][1]);
            }
        }
        const int c0 = 8 * j + 2 * t;
        const double2 hc = *reinterpret_cast<const double2*>(m.hL + c0);
        const double2 sc = *reinterpret_cast<const double2*>(m.sv + c0);
        const double2 hvc = *reinterpret_cast<const double2*>(m.hv + c0);
        const double2 alc = *reinterpret_cast<const double2*>(m.al + c0);
#pragma unroll
        for (int i = j; i < NT; ++i) {
            const int r = 8 * i + g;
            const double hr = m.hL[r], sr = m.sv[r], hvr = m.hv[r], alr = m.al[r];
            double g0 = alr * alc.x - kacc[slot(i, j)][0], g1 = alr * alc.y - kacc[slot(i, j)][1];
            // multiplicity: lower triangle counted twice, diagonal once, (padding / upper part of diagonal tiles) zero
            double w0 = (c0 < r) ? 2.0 : (c0 == r ? 1.0 : 0.0), w1 = (c0 + 1 < r) ? 2.0 : (c0 + 1 == r ? 1.0 : 0.0);
            if (r >= N) w0 = w1 = 0.0;
            g0 *= w0;
            g1 *= w1;
            st_c(m.tiles + slot(i, j) * 64, L, g0, g1);  // weighted G (read by the HF x HF pass)
            if (c0 == r) s_dg += g0;
            if (c0 + 1 == r) s_dg += g1;
            const double t0 = g0 * ((sr * sc.x) * vL * fexp(e[i][0] + (hr + hc.x)));
            const double t1 = g1 * ((sr * sc.y) * vL * fexp(e[i][1] + (hr + hc.y)));
            kacc[slot(i, j)][0] = t0;  // T^L = w G K^L
            kacc[slot(i, j)][1] = t1;
            s_vL += t0 + t1;
            s_rho += t0 * (hvr + hvc.x) + t1 * (hvr + hvc.y);
        }
    }
    {
        const double a0 = warp_sum(s_rho), a1 = warp_sum(s_vL), a2 = warp_sum(s_dg);
        if (lane == 0) {
            m.red[0] = a0;
            m.red[1 + d] = a1;
            m.red[3 + 2 * d] = a2;
        }
    }
    for (int q = 0; q < d; ++q) {
        double sL = 0.0;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const double2 xc = *reinterpret_cast<const double2*>(m.xL + q * NP + 8 * j + 2 * t);
#pragma unroll
            for (int i = j; i < NT; ++i) {
                const double xr = m.xL[q * NP + 8 * i + g];
                const double d0 = xr - xc.x, d1 = xr - xc.y;
                sL = fma(kacc[slot(i, j)][0] * d0, d0, sL);
                sL = fma(kacc[slot(i, j)][1] * d1, d1, sL);
            }
        }
        sL = warp_sum(sL);
        if (lane == 0) m.red[1 + q] = sL;
    }
    __syncwarp();  // weighted G tiles visible to all lanes
    {
        // discrepancy kernel: only HF x HF pairs (h_i h_j = 1).  T^D overwrites G in place.
        const int npairs = nH * (nH + 1) / 2;
        double s_vD = 0.0;
        for (int tt = lane; tt < npairs; tt += 32) {
            int pi = (int)((sqrt(8.0 * tt + 1.0) - 1.0) * 0.5);
            while ((pi + 1) * (pi + 2) / 2 <= tt) ++pi;
            while (pi * (pi + 1) / 2 > tt) --pi;
            const int pj = tt - pi * (pi + 1) / 2;
            const int i = m.hidx[pi], j = m.hidx[pj];  // i >= j
            double ee = m.hD[i] + m.hD[j];
            for (int q = 0; q < d; ++q) ee = fma(m.xD[q * NP + i], m.xD[q * NP + j], ee);
            double* gp = m.tiles + slot(i >> 3, j >> 3) * 64 + tile_off(i & 7, j & 7);
            const double td = (*gp) * vD * fexp(ee);
            *gp = td;
            s_vD += td;
        }
        s_vD = warp_sum(s_vD);
        if (lane == 0) m.red[2 + 2 * d] = s_vD;
        __syncwarp();
        for (int q = 0; q < d; ++q) {
            double sD = 0.0;
            for (int tt = lane; tt < npairs; tt += 32) {
                int pi = (int)((sqrt(8.0 * tt + 1.0) - 1.0) * 0.5);
                while ((pi + 1) * (pi + 2) / 2 <= tt) ++pi;
                while (pi * (pi + 1) / 2 > tt) --pi;
                const int pj = tt - pi * (pi + 1) / 2;
                const int i = m.hidx[pi], j = m.hidx[pj];
                const double df = m.xD[q * NP + i] - m.xD[q * NP + j];
                sD = fma(m.tiles[slot(i >> 3, j >> 3) * 64 + tile_off(i & 7, j & 7)] * df, df, sD);
            }
            sD = warp_sum(sD);
            if (lane == 0) m.red[2 + d + q] = sD;
        }
    }
    __syncwarp();
    for (int q = lane; q < nq && active; q += 32) {
        double f = 1.0;
        if (q == 0) f = 1.0 / rho;
        else if (q <= d) f = m.th[2 * d + 3 + (q - 1)];
        else if (q == d + 1) f = 1.0 / vL;
        else if (q <= 2 * d + 1) f = m.th[3 * d + 3 + (q - d - 2)];
        else if (q == 2 * d + 2) f = 1.0 / vD;
        p.grad[(size_t)prob * nq + q] = -0.5 * f * m.red[q];  // d(nlml) = -1/2 sum G dK
    }
}

template <int NT>
int launch_nt(cudaStream_t st, const SmallArgs& a) {
    const size_t wd = (WarpMem<NT>::doubles(a.d) + 1) & ~(size_t)1;  // keep every warp's base 16-byte aligned
    const size_t bytes = wd * 8 * WPC;
    static int attr_bytes = 0;
    if ((int)bytes > attr_bytes) {
        if (cudaFuncSetAttribute(gpr_small_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
            return -2;
        attr_bytes = (int)bytes;
    }
    gpr_small_mma_kernel<NT><<<(a.B + WPC - 1) / WPC, WPC * 32, bytes, st>>>(a, wd);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace

int launch_gpr_small_mma(cudaStream_t st, const SmallArgs& a) {
    if (a.N < 1 || a.N > 64 || a.d < 1 || a.d > MFGP_SMALL_MAX_D) return -1;
    if (a.B <= 0) return 0;
    switch ((a.N + 7) / 8) {
        case 1: return launch_nt<1>(st, a);
        case 2: return launch_nt<2>(st, a);
        case 3: return launch_nt<3>(st, a);
        case 4: return launch_nt<4>(st, a);
        case 5: return launch_nt<5>(st, a);
        case 6: return launch_nt<6>(st, a);
        case 7: return launch_nt<7>(st, a);
        default: return launch_nt<8>(st, a);
    }
}
