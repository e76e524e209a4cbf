"""CPU model of the DMMA GEMM kernel's data movement (csrc/gemm.cu): TMA boxes with the 128-byte swizzle
-> fragment addresses -> m8n8k4 semantics -> epilogue mapping.  Catches index bugs
without a GPU and proves the shared-memory reads are bank-conflict free."""
import itertools

import numpy as np
import pytest

BK = 16


def frag_row(kmajor, w0, f, g):
    if kmajor:
        return w0 + (f >> 1) * 16 + 2 * g + (f & 1)
    return w0 + (f >> 1) * 16 + 8 * ((g >> 1) & 1) + 4 * (f & 1) + 2 * (g >> 2) + (g & 1)


def frag_addr(kmajor, rows, row, k):
    if kmajor:
        return row * 128 + (((k >> 1) ^ (row & 7)) << 4) + ((k & 1) << 3)
    return (row >> 4) * 2048 + k * 128 + ((((row & 15) >> 1) ^ (k & 7)) << 4) + ((row & 1) << 3)


def swizzle_128b(addr):
    """CU_TENSOR_MAP_SWIZZLE_128B: the 16-byte chunk index (address bits 4-6) is XORed with address bits 7-9; the pattern
    repeats every 1024 bytes, which is why the tile bases are 1024-byte aligned."""
    return addr ^ (((addr >> 7) & 7) << 4)


def tma_box(smem, dst, G, c_inner, c_outer, n_inner, n_outer, inner_is_k, box_outer):
    """One cp.async.bulk.tensor box [box_outer][16 doubles] written densely at `dst` (bytes) and swizzled; coordinates
    outside the tensor (n_inner x n_outer) are zero-filled by the hardware."""
    for o in range(box_outer):
        for i in range(16):
            ci, co = c_inner + i, c_outer + o
            inside = ci < n_inner and co < n_outer
            v = (G(co, ci) if inner_is_k else G(ci, co)) if inside else 0.0
            smem[swizzle_128b(dst + o * 128 + i * 8) // 8] = v


def load_tile(kmajor, rows, G, ld_is_k, row0, nrows, k0, kend):
    """returns smem as array of doubles indexed by byte_addr // 8. G(r, k) global accessor; kend = extent of k."""
    smem = np.full(rows * BK, np.nan)
    if kmajor:
        tma_box(smem, 0, G, k0, row0, kend, nrows, True, rows)
    else:
        for b in range(rows // 16):
            tma_box(smem, b * 2048, G, row0 + 16 * b, k0, nrows, kend, False, BK)
    assert not np.isnan(smem).any()
    return smem


def run_tile(BM, BN, WM, WN, TA, TB, A, B, M, N, K, i0, j0):
    """Emulate one CTA; returns dict {(i,j): value}."""
    a_km, b_km = (not TA), TB
    GA = (lambda r, k: A[k, r]) if TA else (lambda r, k: A[r, k])
    GB = (lambda r, k: B[r, k]) if TB else (lambda r, k: B[k, r])
    nwn = BN // WN
    nwarps = (BM // WM) * nwn
    MT, NTL = WM // 8, WN // 8
    acc = np.zeros((nwarps, 32, MT, NTL, 2))
    nk = (K + BK - 1) // BK
    for it in range(nk):
        sa = load_tile(a_km, BM, GA, None, i0, M, it * BK, K)
        sb = load_tile(b_km, BN, GB, None, j0, N, it * BK, K)
        for warp in range(nwarps):
            wm0, wn0 = (warp // nwn) * WM, (warp % nwn) * WN
            for kk in range(BK // 4):
                # gather fragments for all lanes, then apply MMA semantics
                Af = np.zeros((MT, 8, 4))
                Bf = np.zeros((NTL, 4, 8))
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    k = kk * 4 + t
                    for m in range(MT):
                        Af[m, g, t] = sa[frag_addr(a_km, BM, frag_row(a_km, wm0, m, g), k) // 8]
                    for n in range(NTL):
                        Bf[n, t, g] = sb[frag_addr(b_km, BN, frag_row(b_km, wn0, n, g), k) // 8]
                for m in range(MT):
                    for n in range(NTL):
                        D = Af[m] @ Bf[n]  # 8x8
                        for lane in range(32):
                            g, t = lane >> 2, lane & 3
                            acc[warp, lane, m, n, 0] += D[g, 2 * t]
                            acc[warp, lane, m, n, 1] += D[g, 2 * t + 1]
    out = {}

    def store2(i, j, v0, v1):
        if i >= M or j >= N:
            return
        assert (i, j) not in out
        out[(i, j)] = v0
        if j + 1 < N:
            assert (i, j + 1) not in out
            out[(i, j + 1)] = v1

    for warp in range(nwarps):
        wm0, wn0 = (warp // nwn) * WM, (warp % nwn) * WN
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for m in range(MT):
                i = i0 + frag_row(a_km, wm0, m, g)
                if b_km:
                    for q in range(NTL // 2):
                        j = j0 + wn0 + q * 16 + 4 * t
                        store2(i, j, acc[warp, lane, m, 2 * q, 0], acc[warp, lane, m, 2 * q + 1, 0])
                        store2(i, j + 2, acc[warp, lane, m, 2 * q, 1], acc[warp, lane, m, 2 * q + 1, 1])
                else:
                    for n in range(NTL):
                        store2(i, j0 + frag_row(False, wn0, n, 2 * t), acc[warp, lane, m, n, 0], acc[warp, lane, m, n, 1])
    return out


@pytest.mark.parametrize("TA,TB", list(itertools.product([False, True], repeat=2)))
@pytest.mark.parametrize("cfg", [(64, 64, 32, 32), (128, 128, 64, 32)])
def test_tile_model_matches_matmul(TA, TB, cfg):
    BM, BN, WM, WN = cfg
    rng = np.random.default_rng(1)
    M, N, K = BM - 3, BN - 5, 21  # ragged in every dimension (K odd exercises the 8-byte partial chunk)
    A = rng.standard_normal((K, M) if TA else (M, K))
    B = rng.standard_normal((N, K) if TB else (K, N))
    ref = (A.T if TA else A) @ (B.T if TB else B)
    out = run_tile(BM, BN, WM, WN, TA, TB, A, B, M, N, K, 0, 0)
    assert len(out) == M * N
    got = np.array([[out[(i, j)] for j in range(N)] for i in range(M)])
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("kmajor,rows", [(True, 128), (False, 128), (True, 64), (False, 64)])
def test_fragment_loads_bank_conflict_free(kmajor, rows):
    """64-bit LDS are issued per half-warp: the 16 lanes must touch 16 distinct 8-byte bank pairs."""
    for w0 in range(0, rows, 32):
        for f in range(4):
            for kk in range(4):
                for half in range(2):
                    banks = set()
                    for lane in range(16 * half, 16 * half + 16):
                        g, t = lane >> 2, lane & 3
                        addr = frag_addr(kmajor, rows, frag_row(kmajor, w0, f, g), kk * 4 + t)
                        banks.add((addr // 8) % 16)
                    assert len(banks) == 16, (kmajor, rows, w0, f, kk, half)
