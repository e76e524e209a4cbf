// gpr_small_v5.cu -- K6 v5: batched small-matrix NLML + analytic gradient, TWO warps per problem.
//
// Why (profiles/r01_ncu_gpr_small_v4_final.csv, profiles/r02_fp64_pipe_reconciliation.md): v4 keeps one 53-point problem
// per warp entirely in shared memory, and shared memory (18.6 KB per problem) caps an SM at 12 problems = 12 warps = 3 warps
// per scheduler.  ncu on v4: issue slots 38.7 % busy, stalls per issued instruction `wait` 2.2 (fixed-latency dependency
// chains: DMMA accumulation, the rsqrt / Newton chain of the 8x8 diagonal tiles, the exp polynomial), `math_pipe_throttle`
// 1.7, `short_scoreboard` 1.1 -- the FP64 pipe is saturated in bursts and idle in between because three warps cannot cover
// each other's chains.  Shared memory per problem cannot shrink, so v5 doubles the warps instead: the two warps of a problem
// split every tile loop (even / odd tiles), meet at a 64-thread named barrier where one needs the other's tiles, and run at
// <= 80 registers so that 24 warps (6 per scheduler) are resident.  The 8x8 diagonal-tile factorisation becomes the
// cooperative shuffle form (4 live doubles per lane instead of a private 8x8 copy: 128 instead of 212 FP64 instructions
// per tile); with six warps per scheduler its shuffle latency is covered.
//
// Arithmetic is v4's, operation for operation (same tile products in the same accumulation order, same exp table, same
// pivot handling), so both kernels pass the same parity suite; only the gradient's final cross-warp sums differ in order.
// Algorithm per problem: see gpr_small_v4.cu.  Replaces, per bin, GPR.log_marginal_likelihood + tape.gradient (reference
// mfgpflow/linear.py:206-207).
#include <cmath>
#include <cstdint>

#include "common.cuh"
#include "gpr_small.cuh"
#include "mathx.cuh"

namespace {

constexpr int MAX_SLOTS = 12;  // problems in flight per SM (shared-memory bound), two warps each
constexpr double LOG2PI = 1.8378770664093454835606594728112;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double quad_sum(double v) {  // sum over the 4 lanes that share g
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// the two warps of a problem: named barrier 1 + slot, 64 threads (also orders their shared-memory accesses)
__device__ __forceinline__ void pair_sync(int slot) { asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory"); }

// element offset (doubles) inside a 64-double tile: 16-byte chunks of a row XOR-swizzled by (row & 2)
__host__ __device__ __forceinline__ constexpr int tile_off(int r, int c) {
    return r * 8 + ((((c >> 1) ^ (r & 2))) << 1) + (c & 1);
}
template <int NT>
__device__ __forceinline__ int cslot(int i, int j) {  // i >= j, packed by column
    return j * NT - j * (j - 1) / 2 + (i - j);
}

template <int NT>
struct SlotMem {
    static constexpr int NP = 8 * NT;
    static constexpr int NTRI = NT * (NT + 1) / 2;
    double* tiles;  // [NTRI][64]
    double* xL;     // [d][NP]   inputs / ls_L
    double* hs;     // [NP][2]   { -|xL|^2/2 + log(var_L)/2 , row scale s in {0, 1, rho} }
    double* yv;     // [NP]      y, later alpha
    double* av;     // [NP]      a = W y
    double* inv;    // [2d]      1/ls_L, 1/ls_delta
    double* red;    // [2][2d+4] per-warp partial sums of the gradient, then [0][..] the totals
    double* misc;   // [2]       cross-warp scalars (quadratic form)
    unsigned char* hidx;  // [NP] indices of the HF rows
    __host__ __device__ static size_t doubles(int d) {
        return (size_t)NTRI * 64 + (size_t)d * NP + 4 * NP + 2 * d + 2 * (2 * d + 4) + 2 + NP / 8 + 2;
    }
    __device__ SlotMem(double* base, int d) {
        tiles = base;
        xL = tiles + NTRI * 64;
        hs = xL + d * NP;
        yv = hs + 2 * NP;
        av = yv + NP;
        inv = av + NP;
        red = inv + 2 * d;
        misc = red + 2 * (2 * d + 4);
        hidx = reinterpret_cast<unsigned char*>(misc + 2);
    }
};

// K^L C-fragment of tile (i, j): rows 8i+g, columns 8j+2t, 8j+2t+1 (expanded-square distance folded into the exponent)
template <int NT, int DS>
__device__ __forceinline__ void kl_tile(const SlotMem<NT>& m, const double* etab, int d, int i, int j, int g, int t, double& k0,
                                        double& k1) {
    constexpr int NP = 8 * NT;
    const int r = 8 * i + g, c = 8 * j + 2 * t;
    double e0 = 0.0, e1 = 0.0;
    if constexpr (DS > 0) {
#pragma unroll
        for (int q = 0; q < DS; ++q) {
            const double a = m.xL[q * NP + r];
            const double2 b = *reinterpret_cast<const double2*>(m.xL + q * NP + c);
            e0 = fma(a, b.x, e0);
            e1 = fma(a, b.y, e1);
        }
    } else {
        for (int q = 0; q < d; ++q) {
            const double a = m.xL[q * NP + r];
            const double2 b = *reinterpret_cast<const double2*>(m.xL + q * NP + c);
            e0 = fma(a, b.x, e0);
            e1 = fma(a, b.y, e1);
        }
    }
    const double2 hr = *reinterpret_cast<const double2*>(m.hs + 2 * r);
    const double2 h0 = *reinterpret_cast<const double2*>(m.hs + 2 * c);
    const double2 h1 = *reinterpret_cast<const double2*>(m.hs + 2 * c + 2);
    k0 = (hr.y * h0.y) * fexp_tab(e0 + (hr.x + h0.x), etab);
    k1 = (hr.y * h1.y) * fexp_tab(e1 + (hr.x + h1.x), etab);
}

template <int NT, int DS>
__global__ void __launch_bounds__(MAX_SLOTS * 64, 1) gpr_small_v5_kernel(SmallArgs p, int slot_doubles) {
    constexpr int NP = 8 * NT;
    constexpr int NTRI = NT * (NT + 1) / 2;
    constexpr int HT = (NT + 1) / 2;  // tiles of one block row / column owned by one warp of the pair
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp >> 1, half = warp & 1, oh = 1 - half;
    const int N = p.N, d = (DS > 0) ? DS : p.d;
    const double* etab = smem;  // 2^(j/64), shared by the CTA
    unsigned char* tij = reinterpret_cast<unsigned char*>(smem + 64);  // tile s -> (i << 4) | j, column-packed order
    fexp_table_fill(smem, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) {
        int s = 0;
        for (int j = 0; j < NT; ++j)
            for (int i = j; i < NT; ++i) tij[s++] = (unsigned char)((i << 4) | j);
    }
    __syncthreads();
    const SlotMem<NT> m(smem + 72 + (size_t)slot * slot_doubles, d);
    const int g = lane >> 2, t = lane & 3;
    const int cst = tile_off(g, 2 * t);                        // C fragment: row g, cols 2t, 2t+1
    const int km0 = tile_off(g, t), km1 = tile_off(g, t + 4);  // K-major fragment: (row g, col t + 4s)
    const int mm0 = tile_off(t, g), mm1 = tile_off(t + 4, g);  // M-major fragment: (row t + 4s, col g)
    const int nq = 2 * d + 4;
    const int slots = blockDim.x >> 6;
    double* scr = p.scratch + ((size_t)blockIdx.x * slots + slot) * (NTRI * 64) + 2 * lane;

    for (int prob = blockIdx.x * slots + slot; prob < p.B; prob += gridDim.x * slots) {
        // ---- setup: rows 32 * half .. 32 * half + 31 by this warp ---------------------------------------------------
        const double* __restrict__ theta = p.theta + (size_t)prob * (2 * d + 3);
        if (half == 0 && lane < 2 * d) m.inv[lane] = 1.0 / theta[lane < d ? 1 + lane : 2 + lane];
        const double rho = theta[0], vL = theta[1 + d], vD = theta[2 + 2 * d];
        const double noise = p.noise[prob];
        const double hlv = 0.5 * log(vL);
        pair_sync(slot);
        int nH = 0;
        unsigned long long hfmask = 0ull;  // bit r: row r is a high-fidelity point (both warps build the full mask)
#pragma unroll
        for (int pass = 0; pass < (NP + 31) / 32; ++pass) {
            const int r = lane + 32 * pass;
            bool live = false, hf = false;
            if (r < N) {
                const double fid = p.X[(size_t)r * (d + 1) + d];
                hf = (fid == 1.0);
                live = hf || (fid == 0.0);
            }
            if (r < NP && pass == half) {
                double nL = 0.0;
                for (int q = 0; q < d; ++q) {
                    const double x = live ? p.X[(size_t)r * (d + 1) + q] * m.inv[q] : 0.0;
                    m.xL[q * NP + r] = x;
                    nL = fma(x, x, nL);
                }
                *reinterpret_cast<double2*>(m.hs + 2 * r) = make_double2(fma(-0.5, nL, hlv), live ? (hf ? rho : 1.0) : 0.0);
                m.yv[r] = (r < N) ? p.Y[(size_t)r * p.ldy + (prob + p.prob0) % p.ycols] : 0.0;
            }
            const unsigned hm = __ballot_sync(0xffffffffu, hf);
            if (hf && half == 0) m.hidx[nH + __popc(hm & ((1u << lane) - 1u))] = (unsigned char)r;
            nH += __popc(hm);
            hfmask |= (unsigned long long)hm << (32 * pass);
        }
        pair_sync(slot);

        // ---- 1: covariance tiles; this warp owns the tile pairs {4c + 2 half, 4c + 2 half + 1} --------------------
#pragma unroll 1
        for (int s0 = 2 * half; s0 < NTRI; s0 += 4) {
            const bool two = s0 + 1 < NTRI;
            const int ca = tij[s0], cb = tij[two ? s0 + 1 : s0];
            const int ia = ca >> 4, ja = ca & 15, ib = cb >> 4, jb = cb & 15;
            double ka0, ka1, kb0, kb1;
            kl_tile<NT, DS>(m, etab, d, ia, ja, g, t, ka0, ka1);
            kl_tile<NT, DS>(m, etab, d, ib, jb, g, t, kb0, kb1);
            // K^L is needed again by the gradient: park it in the L2-resident scratch (own lane's values)
            if (p.grad) {
                __stcg(reinterpret_cast<double2*>(scr + s0 * 64), make_double2(ka0, ka1));
                if (two) __stcg(reinterpret_cast<double2*>(scr + (s0 + 1) * 64), make_double2(kb0, kb1));
            }
            if (ia == ja) {  // diagonal: + noise (real rows) or identity (padding rows keep the factorisation well posed)
                const double dg = (8 * ia + g < N) ? noise : 1.0;
                if (g == 2 * t) ka0 += dg;
                if (g == 2 * t + 1) ka1 += dg;
            }
            if (ib == jb) {
                const double dg = (8 * ib + g < N) ? noise : 1.0;
                if (g == 2 * t) kb0 += dg;
                if (g == 2 * t + 1) kb1 += dg;
            }
            *reinterpret_cast<double2*>(m.tiles + s0 * 64 + cst) = make_double2(ka0, ka1);
            if (two) *reinterpret_cast<double2*>(m.tiles + (s0 + 1) * 64 + cst) = make_double2(kb0, kb1);
        }
        pair_sync(slot);
        // discrepancy GP on HF x HF pairs (lower triangle incl. diagonal), straight from the raw inputs
        const int npairs = nH * (nH + 1) / 2;
        if (half == 0) {
            for (int tt = lane; tt < npairs; tt += 32) {
                int pi = (int)((sqrt(8.0 * tt + 1.0) - 1.0) * 0.5);
                while ((pi + 1) * (pi + 2) / 2 <= tt) ++pi;
                while (pi * (pi + 1) / 2 > tt) --pi;
                const int pj = tt - pi * (pi + 1) / 2;
                const int ri = m.hidx[pi], rj = m.hidx[pj];  // ri >= rj
                double ee = 0.0;
                for (int q = 0; q < d; ++q) {
                    const double df = (p.X[(size_t)ri * (d + 1) + q] - p.X[(size_t)rj * (d + 1) + q]) * m.inv[d + q];
                    ee = fma(df, df, ee);
                }
                const double kd = vD * fexp_tab(-0.5 * ee, etab);
                double* tl = m.tiles + cslot<NT>(ri >> 3, rj >> 3) * 64;
                tl[tile_off(ri & 7, rj & 7)] += kd;
                if ((ri >> 3) == (rj >> 3) && ri != rj) tl[tile_off(rj & 7, ri & 7)] += kd;  // diagonal tiles stay symmetric
            }
        }
        pair_sync(slot);

        // ---- 2: left-looking tile Cholesky; warp `half` owns the tiles u = half, half + 2, ... of every column -----
        int bad = 0;
        double lmant = 1.0;  // prod(pivots) = lmant * 2^lexp, renormalised after every diagonal tile (warp 0 of the pair)
        int lexp = 0;
#pragma unroll 1
        for (int kb = 0; kb < NT; ++kb) {
            const int cnt = NT - kb;
            double acc[HT][2];
#pragma unroll
            for (int ul = 0; ul < HT; ++ul) acc[ul][0] = acc[ul][1] = 0.0;
            double* pk = m.tiles + kb * 64;  // tile(kb, 0); tile(kb + u, k) = pk + u * 64
#pragma unroll 1
            for (int k = 0; k < kb; ++k) {
                const double b0 = pk[km0], b1 = pk[km1];
#pragma unroll
                for (int ul = 0; ul < HT; ++ul) {
                    const int u = 2 * ul + half;
                    if (u < cnt) {
                        const double a0 = pk[u * 64 + km0], a1 = pk[u * 64 + km1];
                        dmma(acc[ul][0], acc[ul][1], a0, b0);
                        dmma(acc[ul][0], acc[ul][1], a1, b1);
                    }
                }
                pk += (NT - k - 1) * 64;
            }
            // pk == tile(kb, kb).  acc = K - sum, written back in place
#pragma unroll
            for (int ul = 0; ul < HT; ++ul) {
                const int u = 2 * ul + half;
                if (u < cnt) {
                    const double2 c = *reinterpret_cast<const double2*>(pk + u * 64 + cst);
                    acc[ul][0] = c.x - acc[ul][0];
                    acc[ul][1] = c.y - acc[ul][1];
                    if (u > 0) *reinterpret_cast<double2*>(pk + u * 64 + cst) = make_double2(acc[ul][0], acc[ul][1]);
                }
            }
            __syncwarp();
            // K-major fragments of this warp's raw panel tiles (acc reused as storage); the diagonal tile stays in warp 0's
            // C-fragment registers
            double d0 = acc[0][0], d1 = acc[0][1];
#pragma unroll
            for (int ul = 0; ul < HT; ++ul) {
                const int u = 2 * ul + half;
                if (u > 0 && u < cnt) {
                    acc[ul][0] = pk[u * 64 + km0];
                    acc[ul][1] = pk[u * 64 + km1];
                }
            }
            if (half == 0) {
                // Diagonal tile, cooperatively in its C-fragment layout (lane (g, t) owns A[g][2t], A[g][2t+1] of the SYMMETRIC
                // tile): right-looking 8-step Cholesky with the pivot / row / column values fetched by shuffles, and the
                // inverse built alongside by forward elimination on an identity tile.
                double a0 = d0, a1 = d1;
                double b0 = (g == 2 * t) ? 1.0 : 0.0, b1 = (g == 2 * t + 1) ? 1.0 : 0.0;
                double prod = 1.0;
                int hmin = 0x7fffffff, hmax = 0, first_bad = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const double aj = (j & 1) ? a1 : a0;  // the element of column j this lane may own
                    const double piv = __shfl_sync(0xffffffffu, aj, j * 4 + (j >> 1));
                    const double rv = __shfl_sync(0xffffffffu, aj, (lane & ~3) | (j >> 1));  // A[g][j]
                    const double c0 = __shfl_sync(0xffffffffu, a0, j * 4 + t);              // A[j][2t]   = A[2t][j]
                    const double c1 = __shfl_sync(0xffffffffu, a1, j * 4 + t);              // A[j][2t+1] = A[2t+1][j]
                    const double r0 = __shfl_sync(0xffffffffu, b0, j * 4 + t);              // row j of the inverse so far
                    const double r1 = __shfl_sync(0xffffffffu, b1, j * 4 + t);
                    const int hv = __double2hiint(piv);
                    if ((hv <= 0 || hv >= 0x7ff00000) && !first_bad) first_bad = 8 * kb + j + 1;
                    hmin = min(hmin, hv);
                    hmax = max(hmax, hv);
                    prod *= piv;
                    const double ri = frsqrt(piv);
                    const double lg = rv * ri, lc0 = c0 * ri, lc1 = c1 * ri;
                    a0 = fma(-lg, lc0, a0);  // rows / columns <= j turn into don't-care values that only feed each other
                    a1 = fma(-lg, lc1, a1);
                    const double s0 = r0 * ri, s1 = r1 * ri;
                    const double le = (g > j) ? lg : 0.0;
                    b0 = (g == j) ? s0 : fma(-le, s0, b0);
                    b1 = (g == j) ? s1 : fma(-le, s1, b1);
                }
                if (hmin <= 0 || hmax >= 0x7ff00000) {  // a pivot <= 0 (or denormal), inf or NaN: report the first one
                    if (!bad) bad = first_bad;
                    prod = nan("");
                }
                {
                    const int ph = __double2hiint(prod);
                    lexp += ((ph >> 20) & 0x7ff) - 1023;
                    lmant *= __hiloint2double((ph & 0x800fffff) | 0x3ff00000, __double2loint(prod));
                }
                *reinterpret_cast<double2*>(pk + cst) = make_double2(b0, b1);  // the diagonal slot keeps inv(L_kk)
            }
            pair_sync(slot);
            // panel: L_ik = A_ik inv(L_kk)^T
            {
                const double wb0 = pk[km0], wb1 = pk[km1];
#pragma unroll
                for (int ul = 0; ul < HT; ++ul) {
                    const int u = 2 * ul + half;
                    if (u > 0 && u < cnt) {
                        double c0 = 0.0, c1 = 0.0;
                        dmma(c0, c1, acc[ul][0], wb0);
                        dmma(c0, c1, acc[ul][1], wb1);
                        *reinterpret_cast<double2*>(pk + u * 64 + cst) = make_double2(c0, c1);
                    }
                }
            }
            pair_sync(slot);
        }

        // ---- 3: W = L^-1, row block by row block; this warp owns the columns j = oh, oh + 2, ... (oh = 1 - half:
        //         the column split is mirrored against the Cholesky's so that the warp that factors the diagonal tiles gets the
        //         lighter share here, tests/test_small_pair_schedule_model.py) ----------------
#pragma unroll 1
        for (int i = 1; i < NT; ++i) {
            double acc[HT][2];
#pragma unroll
            for (int jl = 0; jl < HT; ++jl) acc[jl][0] = acc[jl][1] = 0.0;
            const double* li = m.tiles + i * 64;  // tile(i, 0); tile(i, k) = li + offset accumulated below
#pragma unroll 1
            for (int k = 0; k < i; ++k) {
                const double la0 = li[km0], la1 = li[km1];
#pragma unroll
                for (int jl = 0; jl < HT; ++jl) {
                    const int j = 2 * jl + oh;
                    if (j <= k) {
                        const double* wkj = m.tiles + cslot<NT>(k, j) * 64;
                        dmma(acc[jl][0], acc[jl][1], la0, wkj[mm0]);
                        dmma(acc[jl][0], acc[jl][1], la1, wkj[mm1]);
                    }
                }
                li += (NT - k - 1) * 64;  // tile(i, k + 1)
            }
            pair_sync(slot);  // both warps have read row block i of L
#pragma unroll
            for (int jl = 0; jl < HT; ++jl) {
                const int j = 2 * jl + oh;
                if (j < i) *reinterpret_cast<double2*>(m.tiles + cslot<NT>(i, j) * 64 + cst) = make_double2(acc[jl][0], acc[jl][1]);
            }
            __syncwarp();
            const double w0 = -li[km0], w1 = -li[km1];  // li == tile(i, i) = inv(L_ii)
#pragma unroll
            for (int jl = 0; jl < HT; ++jl) {
                const int j = 2 * jl + oh;
                if (j < i) {
                    const double* tl = m.tiles + cslot<NT>(i, j) * 64;
                    acc[jl][0] = acc[jl][1] = 0.0;
                    dmma(acc[jl][0], acc[jl][1], w0, tl[mm0]);
                    dmma(acc[jl][0], acc[jl][1], w1, tl[mm1]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int jl = 0; jl < HT; ++jl) {
                const int j = 2 * jl + oh;
                if (j < i) *reinterpret_cast<double2*>(m.tiles + cslot<NT>(i, j) * 64 + cst) = make_double2(acc[jl][0], acc[jl][1]);
            }
            pair_sync(slot);
        }

        // ---- 4: a = W y (row blocks i = half, half + 2, ...), alpha = W^T a (column blocks likewise), value ----------
        double quad = 0.0;
        {
#pragma unroll 1
            for (int i = half; i < NT; i += 2) {
                double s = 0.0;
                for (int j = 0; j <= i; ++j) {
                    const double2 w = *reinterpret_cast<const double2*>(m.tiles + cslot<NT>(i, j) * 64 + cst);
                    const double2 yy = *reinterpret_cast<const double2*>(m.yv + 8 * j + 2 * t);
                    s = fma(w.x, yy.x, s);
                    s = fma(w.y, yy.y, s);
                }
                s = quad_sum(s);
                if (t == 0) {
                    m.av[8 * i + g] = s;
                    quad = fma(s, s, quad);
                }
            }
            quad = warp_sum(quad);
            if (half == 1 && lane == 0) m.misc[0] = quad;
            pair_sync(slot);  // a complete; y is dead
#pragma unroll 1
            for (int j = half; j < NT; j += 2) {
                double s = 0.0;
                for (int i = j; i < NT; ++i) {
                    const double* tl = m.tiles + cslot<NT>(i, j) * 64;
                    s = fma(tl[mm0], m.av[8 * i + t], s);
                    s = fma(tl[mm1], m.av[8 * i + t + 4], s);
                }
                s = quad_sum(s);
                if (t == 0) m.yv[8 * j + g] = s;  // alpha overwrites y
            }
            if (half == 0 && lane == 0) {
                const double q2 = quad + m.misc[0];
                const double logdet2 = log(lmant) + 0.69314718055994530942 * (double)lexp;  // sum log(pivot) = 2 sum log L_ii
                p.nlml[prob] = 0.5 * q2 + 0.5 * logdet2 + 0.5 * N * LOG2PI;
                if (p.info) p.info[prob] = bad;
                if (bad) atomicCAS(p.d_info, 0, bad);
            }
        }
        pair_sync(slot);  // alpha complete (and the pair leaves the problem together when no gradient is wanted)
        if (!p.grad) continue;

        // ---- 5: K^-1 = W^T W in place (row blocks ascending), G = w o (alpha alpha^T - K^-1); columns j = oh, oh + 2, ... ----
        double s_dg = 0.0;
#pragma unroll 1
        for (int i = 0; i < NT; ++i) {
            double acc[HT][2];
#pragma unroll
            for (int jl = 0; jl < HT; ++jl) acc[jl][0] = acc[jl][1] = 0.0;
            const double* wki = m.tiles + cslot<NT>(i, i) * 64;  // tile(k, i), k = i ...
#pragma unroll 1
            for (int k = i; k < NT; ++k) {
                const double a0 = wki[mm0], a1 = wki[mm1];
#pragma unroll
                for (int jl = 0; jl < HT; ++jl) {
                    const int j = 2 * jl + oh;
                    if (j <= i) {
                        const double* wkj = m.tiles + cslot<NT>(k, j) * 64;
                        const double b0 = (j == i) ? a0 : wkj[mm0];
                        const double b1 = (j == i) ? a1 : wkj[mm1];
                        dmma(acc[jl][0], acc[jl][1], a0, b0);
                        dmma(acc[jl][0], acc[jl][1], a1, b1);
                    }
                }
                wki += 64;
            }
            pair_sync(slot);  // row block i of W is dead from here on (both warps have read tile(i, i))
            const int r = 8 * i + g;
            const double alr = m.yv[r];
#pragma unroll
            for (int jl = 0; jl < HT; ++jl) {
                const int j = 2 * jl + oh;
                if (j <= i) {
                    const int c0 = 8 * j + 2 * t;
                    const double2 alc = *reinterpret_cast<const double2*>(m.yv + c0);
                    double g0 = fma(alr, alc.x, -acc[jl][0]), g1 = fma(alr, alc.y, -acc[jl][1]);
                    // multiplicity: strictly lower counted twice, diagonal once; upper part of diagonal tiles and padding zero
                    double w0 = (c0 < r) ? 2.0 : (c0 == r ? 1.0 : 0.0), w1 = (c0 + 1 < r) ? 2.0 : (c0 + 1 == r ? 1.0 : 0.0);
                    if (r >= N) w0 = w1 = 0.0;
                    g0 *= w0;
                    g1 *= w1;
                    if (j == i) {
                        if (c0 == r) s_dg += g0;
                        if (c0 + 1 == r) s_dg += g1;
                    }
                    *reinterpret_cast<double2*>(m.tiles + cslot<NT>(i, j) * 64 + cst) = make_double2(g0, g1);
                }
            }
        }
        pair_sync(slot);

        // ---- 6: gradient contraction; per-warp partial sums in m.red[half][..], totals by warp 0 ------------------------
        double* red = m.red + half * nq;
        for (int q = lane; q < nq; q += 32) red[q] = 0.0;
        __syncwarp();
        if (half == 0) {  // discrepancy kernel: G at the HF x HF pairs
            for (int t0 = 0; t0 < npairs; t0 += 32) {
                const int tt = t0 + lane;
                double td = 0.0;
                int ri = 0, rj = 0;
                if (tt < npairs) {
                    int pi = (int)((sqrt(8.0 * tt + 1.0) - 1.0) * 0.5);
                    while ((pi + 1) * (pi + 2) / 2 <= tt) ++pi;
                    while (pi * (pi + 1) / 2 > tt) --pi;
                    const int pj = tt - pi * (pi + 1) / 2;
                    ri = m.hidx[pi];
                    rj = m.hidx[pj];
                    double ee = 0.0;
                    for (int q = 0; q < d; ++q) {
                        const double df = (p.X[(size_t)ri * (d + 1) + q] - p.X[(size_t)rj * (d + 1) + q]) * m.inv[d + q];
                        ee = fma(df, df, ee);
                    }
                    td = m.tiles[cslot<NT>(ri >> 3, rj >> 3) * 64 + tile_off(ri & 7, rj & 7)] * vD * fexp_tab(-0.5 * ee, etab);
                }
                const double sv = warp_sum(td);
                if (lane == 0) red[2 + 2 * d] += sv;
                for (int q = 0; q < d; ++q) {
                    const double df = (p.X[(size_t)ri * (d + 1) + q] - p.X[(size_t)rj * (d + 1) + q]) * m.inv[d + q];
                    const double sq = warp_sum(td * df * df);
                    if (lane == 0) red[2 + d + q] += sq;
                }
            }
        }
        if constexpr (DS == 0) pair_sync(slot);  // the generic path overwrites G in place: the HF x HF reads come first
        double s_vL = 0.0, s_rho = 0.0;
        if constexpr (DS > 0) {
            double sL[DS];
#pragma unroll
            for (int q = 0; q < DS; ++q) sL[q] = 0.0;
#pragma unroll 1
            for (int s0 = 2 * half; s0 < NTRI; s0 += 4) {
                const bool two = s0 + 1 < NTRI;
                const int s1 = two ? s0 + 1 : s0;
                const int ca = tij[s0], cb = tij[s1];
                const int ia = ca >> 4, ja = ca & 15, ib = cb >> 4, jb = cb & 15;
                const double2 ka = __ldcg(reinterpret_cast<const double2*>(scr + s0 * 64));
                const double2 kb = __ldcg(reinterpret_cast<const double2*>(scr + s1 * 64));
                const double2 ga = *reinterpret_cast<const double2*>(m.tiles + s0 * 64 + cst);
                double2 gb = *reinterpret_cast<const double2*>(m.tiles + s1 * 64 + cst);
                if (!two) gb.x = gb.y = 0.0;
                const double ta0 = ga.x * ka.x, ta1 = ga.y * ka.y, tb0 = gb.x * kb.x, tb1 = gb.y * kb.y;
                s_vL += (ta0 + ta1) + (tb0 + tb1);
                // exponent of rho in s_i s_j = number of HF points in the pair; only tiles that touch an HF row or column
                if ((((hfmask >> (8 * ia)) | (hfmask >> (8 * ja))) & 0xffull) != 0ull) {
                    const double hr = (double)((hfmask >> (8 * ia + g)) & 1ull);
                    const double hc0 = (double)((hfmask >> (8 * ja + 2 * t)) & 1ull), hc1 = (double)((hfmask >> (8 * ja + 2 * t + 1)) & 1ull);
                    s_rho += ta0 * (hr + hc0) + ta1 * (hr + hc1);
                }
                if ((((hfmask >> (8 * ib)) | (hfmask >> (8 * jb))) & 0xffull) != 0ull) {
                    const double hr = (double)((hfmask >> (8 * ib + g)) & 1ull);
                    const double hc0 = (double)((hfmask >> (8 * jb + 2 * t)) & 1ull), hc1 = (double)((hfmask >> (8 * jb + 2 * t + 1)) & 1ull);
                    s_rho += tb0 * (hr + hc0) + tb1 * (hr + hc1);
                }
                const int ra = 8 * ia + g, cca = 8 * ja + 2 * t, rb = 8 * ib + g, ccb = 8 * jb + 2 * t;
#pragma unroll
                for (int q = 0; q < DS; ++q) {
                    const double xra = m.xL[q * NP + ra], xrb = m.xL[q * NP + rb];
                    const double2 xa = *reinterpret_cast<const double2*>(m.xL + q * NP + cca);
                    const double2 xb = *reinterpret_cast<const double2*>(m.xL + q * NP + ccb);
                    const double da0 = xra - xa.x, da1 = xra - xa.y, db0 = xrb - xb.x, db1 = xrb - xb.y;
                    sL[q] = fma(ta0 * da0, da0, sL[q]);
                    sL[q] = fma(ta1 * da1, da1, sL[q]);
                    sL[q] = fma(tb0 * db0, db0, sL[q]);
                    sL[q] = fma(tb1 * db1, db1, sL[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < DS; ++q) {
                const double v = warp_sum(sL[q]);
                if (lane == 0) red[1 + q] = v;
            }
        } else {
            // generic d: T^L = G o K^L overwrites G (own tiles), then one pass per dimension over the same tiles
#pragma unroll 1
            for (int s0 = 2 * half; s0 < NTRI; s0 += 4) {
#pragma unroll 1
                for (int s = s0; s < s0 + 2 && s < NTRI; ++s) {
                    const int c = tij[s], i = c >> 4, j = c & 15;
                    const double2 kk = __ldcg(reinterpret_cast<const double2*>(scr + s * 64));
                    const double2 gg = *reinterpret_cast<const double2*>(m.tiles + s * 64 + cst);
                    const double t0 = gg.x * kk.x, t1 = gg.y * kk.y;
                    const double hr = (double)((hfmask >> (8 * i + g)) & 1ull);
                    const double hc0 = (double)((hfmask >> (8 * j + 2 * t)) & 1ull), hc1 = (double)((hfmask >> (8 * j + 2 * t + 1)) & 1ull);
                    s_vL += t0 + t1;
                    s_rho += t0 * (hr + hc0) + t1 * (hr + hc1);
                    *reinterpret_cast<double2*>(m.tiles + s * 64 + cst) = make_double2(t0, t1);
                }
            }
            for (int q = 0; q < d; ++q) {
                double sL = 0.0;
#pragma unroll 1
                for (int s0 = 2 * half; s0 < NTRI; s0 += 4) {
#pragma unroll 1
                    for (int s = s0; s < s0 + 2 && s < NTRI; ++s) {
                        const int c = tij[s], i = c >> 4, j = c & 15;
                        const double2 tt = *reinterpret_cast<const double2*>(m.tiles + s * 64 + cst);
                        const double xr = m.xL[q * NP + 8 * i + g];
                        const double2 xc = *reinterpret_cast<const double2*>(m.xL + q * NP + 8 * j + 2 * t);
                        const double d0 = xr - xc.x, d1 = xr - xc.y;
                        sL = fma(tt.x * d0, d0, sL);
                        sL = fma(tt.y * d1, d1, sL);
                    }
                }
                sL = warp_sum(sL);
                if (lane == 0) red[1 + q] = sL;
            }
        }
        {
            const double a0 = warp_sum(s_rho), a1 = warp_sum(s_vL), a2 = warp_sum(s_dg);
            if (lane == 0) {
                red[0] = a0;
                red[1 + d] = a1;
                red[3 + 2 * d] = a2;
            }
        }
        pair_sync(slot);
        if (half == 0) {
            for (int q = lane; q < nq; q += 32) {
                double f = 1.0;
                if (q == 0) f = 1.0 / rho;
                else if (q <= d) f = m.inv[q - 1];
                else if (q == d + 1) f = 1.0 / vL;
                else if (q <= 2 * d + 1) f = m.inv[d + (q - d - 2)];
                else if (q == 2 * d + 2) f = 1.0 / vD;
                p.grad[(size_t)prob * nq + q] = -0.5 * f * (m.red[q] + m.red[nq + q]);  // d(nlml) = -1/2 sum G dK
            }
        }
        pair_sync(slot);  // m.inv / m.red are rewritten by the next problem's setup
    }
}

template <int NT, int DS>
int launch_v5(cudaStream_t st, const SmallArgs& a) {
    const int sd = (int)((SlotMem<NT>::doubles(a.d) + 1) & ~(size_t)1);  // every slot's base stays 16-byte aligned
    const mfgp_dev_info di = mfgp_current_dev_info();
    const int smem_cap = di.smem_optin, sms = di.sms;
    if (smem_cap <= 0 || sms <= 0) return -2;
    int slots = (int)(((size_t)smem_cap - 72 * 8) / ((size_t)sd * 8));
    if (slots > MAX_SLOTS) slots = MAX_SLOTS;
    if (slots < 1) return -1;
    const int want_slots = a.B < sms * slots ? a.B : sms * slots;  // few problems: spread them over the SMs first
    if ((want_slots + sms - 1) / sms < slots) slots = (want_slots + sms - 1) / sms;
    const size_t bytes = (size_t)sd * 8 * slots + 72 * 8;
    static SmemOptIn optin;
    if (!optin.ensure(gpr_small_v5_kernel<NT, DS>, bytes)) return -2;
    const int want = (a.B + slots - 1) / slots;
    const int grid = want < sms ? want : sms;
    SmallArgs b = a;
    b.scratch = nullptr;
    if (a.grad) {  // one K^L slot per problem in flight: <= 148 * 12 * 14 KB = 25 MB, stays in the 126 MB L2
        if (cudaMallocAsync(&b.scratch, (size_t)grid * slots * SlotMem<NT>::NTRI * 64 * sizeof(double), st) != cudaSuccess) return -2;
    }
    gpr_small_v5_kernel<NT, DS><<<grid, slots * 64, bytes, st>>>(b, sd);
    const bool ok = cudaGetLastError() == cudaSuccess;
    if (b.scratch) cudaFreeAsync(b.scratch, st);
    return ok ? 0 : -2;
}

template <int NT>
int launch_nt(cudaStream_t st, const SmallArgs& a) {
    return a.d == 5 ? launch_v5<NT, 5>(st, a) : launch_v5<NT, 0>(st, a);
}

}  // namespace

int launch_gpr_small_v5(cudaStream_t st, const SmallArgs& a) {
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
