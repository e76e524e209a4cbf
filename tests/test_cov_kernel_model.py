"""CPU models of two device-side pieces of the streaming covariance kernels (csrc/cov.cu, csrc/mathx.cuh), so that their
index logic and constants are checked without a GPU:

* fexp512 -- the 512-entry table exp with the integer underflow clamp.  The constants are READ from mathx.cuh; the model
  follows the device code step by step (fma in extended precision) and is compared with a long-double exp.
* tile_dmma -- the rectangular kernel's k augmentation: dot product and both additive exponent terms as d + 2 columns of an
  m8n8k4 DMMA chain, including the zero padding of the last k4 step and the shared-memory pitch that keeps the 8-byte
  fragment loads of a half-warp in 16 different bank pairs."""
import os
import re

import numpy as np
import pytest

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi_fidelity_gpflow_b200", "csrc")


def fexp512_source():
    src = open(os.path.join(CSRC, "mathx.cuh")).read()
    body = src[src.index("__device__ __forceinline__ double fexp512("):]
    return body[: body.index("\n}\n")]


def constants():
    body = fexp512_source()
    num = r"(-?[0-9][0-9.eE+-]*)"
    c = {
        "clamp_hi": int(re.search(r"min\(\(unsigned\)__double2hiint\(x\), (0x[0-9A-Fa-f]+)u\)", body).group(1), 16),
        "scale": float(re.search(r"fma\(x, " + num + r", " + num + r"\)", body).group(1)),
        "magic": float(re.search(r"fma\(x, " + num + r", " + num + r"\)", body).group(2)),
        "ln2_hi": float(re.search(r"double r = fma\(nf, " + num + r", x\)", body).group(1)),
        "ln2_lo": float(re.search(r"r = fma\(nf, " + num + r", r\)", body).group(1)),
        "c4": float(re.search(r"double q = fma\(r, " + num + r", " + num + r"\)", body).group(1)),
        "c3": float(re.search(r"double q = fma\(r, " + num + r", " + num + r"\)", body).group(2)),
    }
    assert "tab[n & 511]" in body and "(n >> 9) << 20" in body
    return c


def fexp512_model(x, c):
    L = np.longdouble
    fma = lambda a, b, d: (L(a) * L(b) + L(d)).astype(np.float64)
    x = np.asarray(x, dtype=np.float64)
    bits = x.view(np.uint64)
    hi = np.minimum((bits >> np.uint64(32)).astype(np.uint32), np.uint32(c["clamp_hi"]))
    x = ((hi.astype(np.uint64) << np.uint64(32)) | (bits & np.uint64(0xFFFFFFFF))).view(np.float64)
    t = fma(x, c["scale"], c["magic"])
    n = (t.view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32).astype(np.int32)
    nf = t - c["magic"]
    r = fma(nf, c["ln2_hi"], x)
    r = fma(nf, c["ln2_lo"], r)
    T = np.exp2((n & 511) / 512.0)
    q = fma(r, c["c4"], c["c3"])
    q = fma(q, r, 0.5)
    p = fma(q, r * r, r)
    v = fma(T, p, T)
    out = (v.view(np.uint64).astype(np.int64) + ((n.astype(np.int64) >> 9) << 52)).view(np.float64)
    return out, r


def test_fexp512_constants():
    c = constants()
    assert c["clamp_hi"] == np.float64(-700.0).view(np.uint64) >> np.uint64(32)
    assert c["scale"] == 512.0 / np.log(2.0) and c["magic"] == 1.5 * 2.0**52
    # Cody-Waite split of ln2/512: the high part has >= 20 trailing zero bits (n has < 2^20 magnitude for x >= -700, so
    # n * hi is exact), hi + lo reproduces ln2/512 to ~1e-32
    hi, lo = -c["ln2_hi"], -c["ln2_lo"]
    assert (np.float64(hi).view(np.uint64) & np.uint64((1 << 20) - 1)) == 0
    assert abs((np.longdouble(hi) + np.longdouble(lo)) - np.log(np.longdouble(2)) / 512) < 1e-21
    assert 700.0 * c["scale"] < 2**20
    assert c["c4"] == 1.0 / 24 and c["c3"] == 1.0 / 6


def test_fexp512_accuracy_and_clamp():
    c = constants()
    rng = np.random.default_rng(0)
    x = np.concatenate([-60 * rng.random(400000), 3 * rng.random(20000), -700 * rng.random(50000),
                        [0.0, -0.0, -1e-300, 1e-17, -699.999, -700.0]])
    y, r = fexp512_model(x, c)
    ref = np.exp(x.astype(np.longdouble))
    rel = np.abs((y.astype(np.longdouble) - ref) / ref).astype(np.float64)
    assert np.abs(r).max() <= np.log(2) / 1024 * 1.001  # round-to-nearest of x * 512/ln2 (itself rounded) to an integer
    assert rel.max() < 1.5 * 2.0**-53 * 2  # < 1.5 ulp (the GPU adds nothing: every step above is one fma)
    # far below -700, incl. products x * 512/ln2 beyond the int32 range: clamped to exp(-700.x), finite, tiny, never garbage
    far = np.array([-700.5, -5000.0, -1e9, -1e300, -np.inf])
    yf, _ = fexp512_model(far, c)
    assert np.all(np.isfinite(yf)) and np.all(yf > 0) and np.all(yf < 1.1e-304)
    assert np.isnan(fexp512_model(np.array([np.nan]), c)[0][0])


@pytest.mark.parametrize("d", [1, 2, 5, 6, 7, 10, 16])
def test_tile_dmma_k_augmentation(d):
    """Lane (g, t) of k4 step ks supplies A[row g][k = 4 ks + t] and B[k][col g] (cov.cu: tile_dmma).  k < d: coordinates;
    k = d: (hA, 1); k = d + 1: (1, hB); beyond: zeros.  The chain must give a.b + hA + hB for every row/column pair."""
    rng = np.random.default_rng(d)
    a, b = rng.standard_normal((8, d)), rng.standard_normal((8, d))
    hA, hB = rng.standard_normal(8), rng.standard_normal(8)
    KS = (d + 2 + 3) // 4
    acc = np.zeros((8, 8))
    for ks in range(KS):
        Af, Bf = np.zeros((8, 4)), np.zeros((4, 8))
        for t in range(4):
            k = 4 * ks + t
            la, lb = k <= d, (k < d) or (k == d + 1)
            for g in range(8):
                Af[g, t] = (a[g, k] if k < d else hA[g]) if la else (1.0 if k == d + 1 else 0.0)
                Bf[t, g] = (b[g, k] if k < d else hB[g]) if lb else (1.0 if k == d else 0.0)
        acc += Af @ Bf  # one m8n8k4
    np.testing.assert_allclose(acc, a @ b.T + hA[:, None] + hB[None, :], rtol=1e-13, atol=1e-13)
    assert 4 * KS >= d + 2 and 4 * (KS - 1) < d + 2


def test_rect_kernel_fragment_loads_are_conflict_free():
    """Panel rows are padded to COV_PS doubles; a fragment load reads address (k row) * COV_PS + base + g with k = 4 ks + t.
    Shared memory serves an 8-byte access per half-warp: its 16 lanes must fall into 16 different bank pairs."""
    src = open(os.path.join(CSRC, "cov.cu")).read()
    ps = int(re.search(r"constexpr int COV_PS = (\d+);", src).group(1))
    assert ps % 2 == 0 and ps >= 64  # 16-byte aligned rows for cp.async, room for 64 points
    for ks in range(3):
        for half in range(2):
            lanes = range(16 * half, 16 * half + 16)
            banks = {((4 * ks + (lane & 3)) * ps + (lane >> 2)) % 16 for lane in lanes}
            assert len(banks) == 16
