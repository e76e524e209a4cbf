"""GPU parity tests of the CUDA building blocks (through the C-ABI) against the CPU oracle."""
import numpy as np
import pytest

from oracle import mfgp_oracle as onp
from oracle import mfgp_oracle_torch as otc
from tests._helpers import goldens

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h():
    from multi_fidelity_gpflow_b200 import _lib

    return _lib.Handle(0)


def rand_theta(rng, d):
    return onp.pack_theta(rng.uniform(0.5, 2.0), rng.uniform(0.3, 2.0, d), rng.uniform(0.5, 2.0),
                          rng.uniform(0.3, 2.0, d), rng.uniform(0.2, 1.5))


def rand_X(rng, n, d, frac_h=0.3, shuffle=True):
    x = rng.random((n, d))
    f = (rng.random(n) < frac_h).astype(float)
    if not shuffle:
        f = np.sort(f)
    return np.hstack([x, f[:, None]])


# ------------------------------------------------------------------------------------ K1 cov
@pytest.mark.parametrize("N,N2,d", [(53, 10, 5), (1, 1, 1), (64, 64, 3), (65, 129, 10), (300, 1164, 10), (200, 77, 16)])
def test_cov_rect(h, N, N2, d):
    rng = np.random.default_rng(N * 1000 + N2)
    X, X2, th = rand_X(rng, N, d), rand_X(rng, N2, d), rand_theta(rng, d)
    K = h.cov(X, X2, th)
    np.testing.assert_allclose(K, onp.mf_K(X, X2, th), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("N,d", [(53, 5), (80, 1), (129, 10), (1164, 10), (2, 2)])
def test_cov_symmetric_mirror(h, N, d):
    rng = np.random.default_rng(N)
    X, th = rand_X(rng, N, d), rand_theta(rng, d)
    K = h.cov(X, None, th)
    ref = onp.mf_K(X, None, th)
    np.testing.assert_allclose(K, ref, rtol=1e-12, atol=1e-14)
    assert np.array_equal(K, K.T)  # mirrored store: exactly symmetric


def test_cov_dead_rows_and_nan_fidelity(h):
    """Quirk Q1: fidelity not exactly 0/1 (0.5, 1-ulp, NaN) -> zero rows/cols (linear.py:82)."""
    rng = np.random.default_rng(7)
    X = rand_X(rng, 70, 4)
    X[3, -1] = 0.5
    X[10, -1] = np.nextafter(1.0, 0.0)
    X[20, -1] = np.nan
    X[20, 0] = np.nan
    th = rand_theta(rng, 4)
    K = h.cov(X, None, th)
    ref = onp.mf_K(X, None, th)
    assert np.all(K[[3, 10, 20]] == 0) and np.all(K[:, [3, 10, 20]] == 0)
    np.testing.assert_allclose(K, ref, rtol=1e-12, atol=1e-14)
    np.testing.assert_array_equal(h.cov_diag(X, th), onp.mf_K_diag(X, th))


def test_cov_golden_datasets(h):
    for name, d in (("hbs", 5), ("goku", 10)):
        ds = onp.load_dataset(name)
        th = onp.default_theta(d)
        np.testing.assert_allclose(h.cov(ds["X"], None, th), onp.mf_K(ds["X"], None, th), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(h.cov(ds["X"], ds["X_test"], th), onp.mf_K(ds["X"], ds["X_test"], th), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(h.cov_diag(ds["X_test"], th), onp.mf_K_diag(ds["X_test"], th), rtol=1e-15)


# ------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (53, 49, 53), (300, 1164, 300), (257, 130, 19), (2, 2, 2), (1000, 8, 1000), (64, 640, 16)])
def test_gemm(h, ta, tb, m, n, k):
    rng = np.random.default_rng(m + n + k)
    ev = lambda x: x + (x & 1)
    A = np.zeros((k, ev(m)) if ta else (m, ev(k)))
    B = np.zeros((n, ev(k)) if tb else (k, ev(n)))
    A[:, : (m if ta else k)] = rng.standard_normal((k, m) if ta else (m, k))
    B[:, : (k if tb else n)] = rng.standard_normal((n, k) if tb else (k, n))
    opA = (A.T if ta else A)[:m, :k]
    opB = (B.T if tb else B)[:k, :n]
    C0 = rng.standard_normal((m, n))
    # the binding derives m/n/k from shapes: pass exact-shape views through padded buffers
    from multi_fidelity_gpflow_b200 import _lib
    import ctypes as C

    out = C0.copy()
    rc = _lib._lib.mfgp_gemm(h._h, b"T" if ta else b"N", b"T" if tb else b"N", m, n, k, 0.7, _lib._ptr(A), A.shape[1],
                             _lib._ptr(B), B.shape[1], -0.3, _lib._ptr(out), n)
    assert rc == 0, _lib._lib.mfgp_last_error(h._h)
    ref = 0.7 * opA @ opB - 0.3 * C0
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-11)


# ------------------------------------------------------------------------------------ potrf / trtri
@pytest.mark.parametrize("N", [1, 5, 53, 128, 129, 300, 640, 1164])
def test_potrf_and_inverse(h, N):
    rng = np.random.default_rng(N)
    d = 4
    X, th = rand_X(rng, N, d), rand_theta(rng, d)
    K = onp.mf_K(X, None, th) + 1e-2 * np.eye(N)
    lda = N + (N & 1)
    A = np.zeros((N, lda))
    A[:, :N] = np.tril(K)
    L, W = h.potrf(A, want_inverse=True)
    Lr = np.linalg.cholesky(K)
    np.testing.assert_allclose(L[:, :N], Lr, rtol=1e-9, atol=1e-11)
    assert np.all(np.triu(L[:, :N], 1) == 0)
    Wr = np.linalg.inv(Lr)
    np.testing.assert_allclose(W[:, :N], Wr, rtol=1e-8, atol=1e-8 * np.abs(Wr).max())
    np.testing.assert_allclose(W[:, :N] @ Lr, np.eye(N), atol=1e-8)


def test_potrf_and_inverse_large_outer_block_512(h):
    """N = 13 000 is just past the switch to rank-512 trailing updates (csrc/chol.cu, N > 12 288) and exercises the
    look-ahead / priority-stream chain that bench.py times at N = 16 384; checked against LAPACK."""
    N = 13000
    ds = onp.synthetic_exact_dataset(N)
    K = onp.mf_K(ds["X"], None, ds["theta"])
    K[np.diag_indices(N)] += ds["noise"]
    Lr = np.linalg.cholesky(K)
    A = np.tril(K)
    del K
    L, W = h.potrf(A, want_inverse=True)
    del A
    # conditioning ~1e3 N: compare in the backward-error sense the factorisation guarantees
    scale = np.abs(Lr).max()
    assert np.abs(L - Lr).max() < 1e-10 * scale
    assert np.all(np.triu(L[:256, :256], 1) == 0)
    # W = inv(L): check W L = I on a slab of rows (the full product would need a 13 000^3 CPU GEMM; 1/13 of it is enough
    # to cover every merge level of the blocked trtri)
    rows = np.r_[0:300, 6400:6700, N - 400:N]
    R = W[rows] @ Lr
    R[np.arange(rows.size), rows] -= 1.0
    assert np.abs(R).max() < 1e-8


@pytest.mark.parametrize("N", [4096, 13000])
def test_gpr_nlml_grad_large_vs_oracle(h, N):
    """Exact-GPR objective + gradient beyond the sizes of the datasets: N = 4096 (rank-256 path) and N = 13 000 (rank-512
    path), C5 synthetic data, against the oracle's closed form (validated against autograd + finite differences in
    tests/test_oracle_fd.py).  North-star tolerances: 1e-9 on the value, 1e-7 on the gradient."""
    ds = onp.synthetic_exact_dataset(N)
    nlml, g = h.gpr_nlml_grad(ds["X"], ds["Y"], ds["theta"], ds["noise"])
    lml, gth, gnz = onp.gpr_lml_grad_analytic(ds["X"], ds["Y"], ds["theta"], ds["noise"])
    assert abs(nlml + lml) < 1e-9 * abs(lml), (nlml, lml)
    ref = -np.concatenate([gth, [gnz]])
    np.testing.assert_allclose(g, ref, rtol=1e-7, atol=1e-7 * np.abs(ref).max())
    assert abs(h.gpr_nlml(ds["X"], ds["Y"], ds["theta"], ds["noise"]) - nlml) < 1e-12 * abs(nlml)


@pytest.mark.parametrize("N,P", [(300, 1), (1164, 3), (129, 64)])
def test_gpr_value_only_path_matches_gradient_path(h, N, P):
    """Value-only evaluations skip the triangular inverse (a = L^-1 Y by blocked forward substitution with the
    diagonal-block inverses); the value must equal the one the NLML + gradient path computes, single and batched,
    ragged last block and P columns included."""
    rng = np.random.default_rng(N + P)
    X, th = rand_X(rng, N, 4), rand_theta(rng, 4)
    Y = rng.standard_normal((N, P))
    v = h.gpr_nlml(X, Y, th, 1e-2)
    vg, _ = h.gpr_nlml_grad(X, Y, th, 1e-2)
    assert abs(v - vg) < 1e-11 * abs(vg)
    assert abs(v + onp.gpr_lml(X, Y, th, 1e-2)) < 1e-9 * abs(v)
    ths = np.stack([rand_theta(rng, 4) for _ in range(5)])
    nz = np.full(5, 1e-2)
    Yb = rng.standard_normal((N, 5))
    vb, _ = h.gpr_batched_nlml_grad(X, Yb, ths, nz, want_grad=False)
    vbg, _ = h.gpr_batched_nlml_grad(X, Yb, ths, nz)
    np.testing.assert_allclose(vb, vbg, rtol=1e-11)


def test_potrf_not_positive_definite(h):
    from multi_fidelity_gpflow_b200._lib import NotPositiveDefiniteError

    A = np.eye(200)
    A[150, 150] = -1.0
    with pytest.raises(NotPositiveDefiniteError) as e:
        h.potrf(A)
    assert "151" in str(e.value)
    h.potrf(np.eye(10))  # the handle recovers


# ------------------------------------------------------------------------------------ exact GPR
def test_gpr_golden_G1_G2(h):
    G = goldens()
    for name, d, key in (("hbs", 5, "G1_hbs_gpr_lml_init"), ("goku", 10, "G2_goku_gpr_lml_init")):
        ds = onp.load_dataset(name)
        nlml = h.gpr_nlml(ds["X"], ds["Y"], onp.default_theta(d), 1e-3)
        assert abs(-nlml - G[key]["value"]) / abs(G[key]["value"]) < 1e-9  # north-star tolerance


@pytest.mark.parametrize("name,d", [("hbs", 5), ("goku", 10), ("forrester", 1)])
def test_gpr_nlml_grad_vs_oracle(h, name, d):
    rng = np.random.default_rng(3)
    ds = onp.forrester_dataset() if name == "forrester" else onp.load_dataset(name)
    X, Y = ds["X"], ds["Y"]
    for th in (onp.default_theta(d), rand_theta(rng, d)):
        nlml, g = h.gpr_nlml_grad(X, Y, th, 1e-3)
        lml, gth, gnz = otc.gpr_lml_value_and_grad(X, Y, th, 1e-3)
        assert abs(nlml + lml) / abs(lml) < 1e-9
        ref = -np.concatenate([gth, [gnz]])
        np.testing.assert_allclose(g, ref, rtol=1e-7, atol=1e-7 * np.abs(ref).max())


def test_gpr_predict_vs_oracle(h):
    for name, d in (("hbs", 5), ("goku", 10)):
        ds = onp.load_dataset(name)
        th = onp.default_theta(d)
        mean, var = h.gpr_predict(ds["X"], ds["Y"], ds["X_test"], th, 1e-3)
        rm, rv = onp.gpr_predict(ds["X"], ds["Y"], ds["X_test"], th, 1e-3)
        np.testing.assert_allclose(mean, rm, rtol=1e-8, atol=1e-8 * np.abs(rm).max())
        np.testing.assert_allclose(var, rv, rtol=1e-8, atol=1e-8 * np.abs(rv).max())
    ds = onp.forrester_dataset()
    th = onp.default_theta(1)
    for Xp in (ds["X_plot_L"], ds["X_plot_H"]):
        mean, var = h.gpr_predict(ds["X"], ds["Y"], Xp, th, 1e-3)
        rm, rv = onp.gpr_predict(ds["X"], ds["Y"], Xp, th, 1e-3)
        np.testing.assert_allclose(mean, rm, rtol=1e-8, atol=1e-8 * np.abs(rm).max())
        np.testing.assert_allclose(var, rv, rtol=1e-8, atol=1e-8 * np.abs(rv).max())


@pytest.mark.parametrize("Ns", [3, 400, 1700])
def test_gpr_predict_both_solve_paths(h, Ns):
    """N = 1500: up to N test points go through the forward substitution A_s = L^-1 K_s, more through W = L^-1 and a
    GEMM; both against the oracle (north-star tolerance 1e-8), P = 2 output columns, ragged last block."""
    rng = np.random.default_rng(Ns)
    X, Xs, th = rand_X(rng, 1500, 3), rand_X(rng, Ns, 3), rand_theta(rng, 3)
    Y = rng.standard_normal((1500, 2))
    mean, var = h.gpr_predict(X, Y, Xs, th, 1e-2)
    rm, rv = onp.gpr_predict(X, Y, Xs, th, 1e-2)
    np.testing.assert_allclose(mean, rm, rtol=1e-8, atol=1e-8 * np.abs(rm).max())
    np.testing.assert_allclose(var, rv, rtol=1e-8, atol=1e-8 * np.abs(rv).max())


# ------------------------------------------------------------------------------------ batched per-bin
def test_batched_small_hbs_vs_oracle(h):
    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"]
    B, d = Y.shape[1], 5
    rng = np.random.default_rng(0)
    th = np.tile(onp.default_theta(d), (B, 1)) * np.exp(0.3 * rng.standard_normal((B, 2 * d + 3)))
    nz = np.full(B, 1e-3) * np.exp(0.3 * rng.standard_normal(B))
    nlml, grad = h.gpr_batched_nlml_grad(X, Y, th, nz)
    vals, grads = otc.gpr_batched_value_and_grad(X, Y, th, nz)
    np.testing.assert_allclose(nlml, -vals, rtol=1e-9)
    for b in range(B):
        np.testing.assert_allclose(grad[b], -grads[b], rtol=1e-7, atol=1e-7 * np.abs(grads[b]).max())


def test_batched_identity_sum_equals_shared_and_G1(h):
    ds = onp.load_dataset("hbs")
    th = np.tile(onp.default_theta(5), (49, 1))
    nlml, _ = h.gpr_batched_nlml_grad(ds["X"], ds["Y"], th, np.full(49, 1e-3))
    g1 = goldens()["G1_hbs_gpr_lml_init"]["value"]
    assert abs(-nlml.sum() - g1) / abs(g1) < 1e-9
    shared = h.gpr_nlml(ds["X"], ds["Y"], onp.default_theta(5), 1e-3)
    assert abs(nlml.sum() - shared) / abs(shared) < 1e-10


@pytest.mark.parametrize("N,d,frac", [(1, 1, 0.0), (2, 3, 1.0), (17, 2, 0.5), (64, 16, 0.3), (33, 7, 0.0)])
def test_batched_small_edge_shapes(h, N, d, frac):
    rng = np.random.default_rng(N)
    X = rand_X(rng, N, d, frac)
    B = 5
    Y = rng.standard_normal((N, B))
    th = np.stack([rand_theta(rng, d) for _ in range(B)])
    nz = rng.uniform(1e-3, 1e-1, B)
    nlml, grad = h.gpr_batched_nlml_grad(X, Y, th, nz)
    vals, grads = otc.gpr_batched_value_and_grad(X, Y, th, nz)
    np.testing.assert_allclose(nlml, -vals, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(grad, -grads, rtol=1e-7, atol=1e-7 * max(1.0, np.abs(grads).max()))


def test_batched_blocked_path_forrester(h):
    """N = 80 > 64 routes through the blocked (potrf/trtri/GEMM) batched path."""
    ds = onp.forrester_dataset()
    rng = np.random.default_rng(5)
    B = 4
    Y = np.hstack([ds["Y"] + 0.1 * rng.standard_normal(ds["Y"].shape) for _ in range(B)])
    th = np.stack([rand_theta(rng, 1) for _ in range(B)])
    nz = np.full(B, 1e-3)
    nlml, grad = h.gpr_batched_nlml_grad(ds["X"], Y, th, nz)
    vals, grads = otc.gpr_batched_value_and_grad(ds["X"], Y, th, nz)
    np.testing.assert_allclose(nlml, -vals, rtol=1e-9)
    np.testing.assert_allclose(grad, -grads, rtol=1e-7, atol=1e-7 * np.abs(grads).max())


def test_batched_not_pd_reports_per_problem_info(h):
    from multi_fidelity_gpflow_b200._lib import NotPositiveDefiniteError

    rng = np.random.default_rng(1)
    X = rand_X(rng, 20, 2, 0.0)
    X[5] = X[4]  # duplicate point + negative noise -> not PD for problem 1 only
    Y = rng.standard_normal((20, 3))
    th = np.tile(onp.default_theta(2), (3, 1))
    nz = np.array([1e-3, -1e-3, 1e-3])
    info = np.zeros(3, dtype=np.int32)
    nl, _ = h.gpr_batched_nlml_grad(X, Y, th, nz, info=info)  # per-problem status: no exception
    assert info[0] == 0 and info[2] == 0 and info[1] > 0
    assert np.isfinite(nl[[0, 2]]).all() and not np.isfinite(nl[1])
    with pytest.raises(NotPositiveDefiniteError):  # without info[] the failure is the call's error
        h.gpr_batched_nlml_grad(X, Y, th, nz)


def test_cov_streaming_kernel_full_size_properties(h):
    """BASELINE config 5 size (N = 16 384, d = 10, 1/8 HF): the streaming K1 path (prescale + persistent CTAs + bulk-store
    mirror).  Size-independent properties: K(X,X) is exactly symmetric, equals the rectangular evaluation K(X, X2 = copy)
    to rounding, has var_L / rho^2 var_L + var_delta on the diagonal, and sampled elements match the oracle to 1e-12."""
    import torch

    from multi_fidelity_gpflow_b200 import _lib

    ds = onp.synthetic_exact_dataset(16384)
    X, th = ds["X"], ds["theta"]
    N, d = X.shape[0], X.shape[1] - 1
    dev = torch.device("cuda:0")
    tX, tth = torch.from_numpy(X).to(dev), torch.from_numpy(th).to(dev)
    K = torch.empty(N, N, dtype=torch.float64, device=dev)
    assert _lib._lib.mfgp_cov(h._h, _lib._ptr(tX), N, None, N, d, _lib._ptr(tth), _lib._ptr(K), N) == 0
    torch.cuda.synchronize()
    assert torch.equal(K, K.T)
    tX2 = tX.clone()
    K2 = torch.empty(N, N, dtype=torch.float64, device=dev)
    assert _lib._lib.mfgp_cov(h._h, _lib._ptr(tX), N, _lib._ptr(tX2), N, d, _lib._ptr(tth), _lib._ptr(K2), N) == 0
    torch.cuda.synchronize()
    assert float((K - K2).abs().max()) <= 1e-13 * float(K.abs().max())
    np.testing.assert_allclose(K.diagonal().cpu().numpy(), onp.mf_K_diag(X, th), rtol=1e-13)
    rng = np.random.default_rng(0)
    ii, jj = rng.integers(0, N, 4000), rng.integers(0, N, 4000)
    ii[:500] = rng.integers(N * 7 // 8, N, 500)  # make sure HF x HF pairs are sampled
    jj[:500] = rng.integers(N * 7 // 8, N, 500)
    got = K[torch.from_numpy(ii).to(dev), torch.from_numpy(jj).to(dev)].cpu().numpy()
    ref = np.array([onp.mf_K(X[i:i + 1], X[j:j + 1], th)[0, 0] for i, j in zip(ii, jj)])
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("d", [5, 7, 10])
def test_cov_streaming_kernel_vs_oracle(h, d):
    """The streaming K1 path (taken from 1024 tiles on) element by element against the oracle: d = 5 and 10 are the compiled-in
    dimensions, d = 7 the run-time one; ragged sizes, shuffled fidelities, rho != 1."""
    rng = np.random.default_rng(40 + d)
    X, X2, th = rand_X(rng, 2101, d), rand_X(rng, 2075, d), rand_theta(rng, d)
    np.testing.assert_allclose(h.cov(X, X2, th), onp.mf_K(X, X2, th), rtol=1e-12, atol=1e-14)
    Xs = rand_X(rng, 2950, d)
    K = h.cov(Xs, None, th)
    np.testing.assert_allclose(K, onp.mf_K(Xs, None, th), rtol=1e-12, atol=1e-14)
    assert np.array_equal(K, K.T)


def test_cov_streaming_kernel_far_points_underflow(h):
    """Tiny length-scales: exponents far below -700 (and beyond the int32 range of the table index) must give |K| < 1e-300
    off the diagonal, not garbage.  (The diagonal carries the cancellation error of the expanded squared distance the
    reference uses too, eps * |x / ls|^2 in the exponent, hence the loose tolerance there.)"""
    d = 5
    rng = np.random.default_rng(3)
    X, th = rand_X(rng, 2950, d), rand_theta(rng, d)
    th[1:1 + d] = 1e-4   # kernel_L length-scales
    th[2 + d:2 + 2 * d] = 1e-5
    K = h.cov(X, None, th)
    off = K[~np.eye(len(X), dtype=bool)]
    assert np.all(np.isfinite(K)) and np.all(off >= 0.0) and off.max() < 1e-300
    np.testing.assert_allclose(np.diag(K), onp.mf_K_diag(X, th), rtol=1e-4)
    K2 = h.cov(X[:2101], X[100:2175], th)
    assert np.all(np.isfinite(K2)) and np.all(K2 >= 0.0)
    i, j = np.arange(100, 2101), np.arange(0, 2001)  # the shared points sit on this diagonal of the block
    assert np.delete(K2.ravel(), i * K2.shape[1] + j).max() < 1e-300


@pytest.mark.parametrize("N,d", [(53, 5), (130, 3), (700, 10)])
def test_cov_grad_contraction_vs_autograd(h, N, d):
    """mfgp_cov_grad: sum_ij G_ij dK_ij/dtheta for a symmetric G given by its lower triangle (the backward pass of K)."""
    import torch

    rng = np.random.default_rng(N)
    X = rand_X(rng, N, d, 0.3)
    th = rand_theta(rng, d)
    A = rng.standard_normal((N, N))
    G = A + A.T
    Glow = np.tril(G) + np.triu(rng.standard_normal((N, N)), 1)  # garbage above the diagonal must be ignored
    Glow = np.ascontiguousarray(np.pad(Glow, ((0, 0), (0, N % 2))))  # even leading dimension
    got = h.cov_grad(X, th, Glow, scale=-0.5)
    tht = torch.tensor(th, requires_grad=True)
    K = otc.mf_K(torch.from_numpy(X), torch.from_numpy(X), tht)
    (torch.from_numpy(G) * K).sum().backward()
    ref = -0.5 * np.concatenate([tht.grad.numpy(), [np.trace(G)]])
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())


def test_dlpack_device_exporter_zero_copy(h):
    """A CUDA tensor that only speaks DLPack (as a TF / JAX tensor would) crosses the C-ABI as a device pointer."""
    import torch

    from multi_fidelity_gpflow_b200 import _lib

    class OnlyDLPack:
        def __init__(self, t):
            self._t, self.shape = t, tuple(t.shape)

        def __dlpack__(self, stream=None):
            return self._t.__dlpack__()

    rng = np.random.default_rng(0)
    X, th = rand_X(rng, 300, 5), rand_theta(rng, 5)
    Xd, thd = torch.from_numpy(X).cuda(), torch.from_numpy(th).cuda()
    Kd = torch.zeros(300, 300, dtype=torch.float64, device="cuda")
    addr, shape, kind = _lib.from_dlpack_capsule(Xd.__dlpack__())
    assert addr == Xd.data_ptr() and shape == (300, 6) and kind == "cuda"
    h.cov(OnlyDLPack(Xd), None, OnlyDLPack(thd), out=OnlyDLPack(Kd))
    assert h.sync() == 0
    np.testing.assert_array_equal(Kd.cpu().numpy(), h.cov(X, None, th))


def test_peer_store_one_to_many(h):
    """mfgp_peer_store: one kernel stores a block into several destination buffers (here local ones; in dist_chol.py they
    are the peers' symmetric-memory buffers over NVLink -- covered by tests/test_multi_gpu.py on a multi-GPU box)."""
    import torch

    src = torch.randn(1024, 130, dtype=torch.float64, device="cuda")
    dsts = [torch.zeros_like(src) for _ in range(5)]
    h.peer_store(src, [t.data_ptr() for t in dsts], src.numel())
    assert h.sync() == 0
    for t in dsts:
        assert torch.equal(t, src)
    with pytest.raises(ValueError):
        h.peer_store(src, [t.data_ptr() for t in dsts], src.numel() - 1)  # odd count
    with pytest.raises(ValueError):
        h.peer_store(src, [dsts[0].data_ptr()] * 9, src.numel())  # > MFGP_PEER_MAX destinations


def test_graph_mem_trim_after_captured_library_calls(h):
    """Library calls captured into a caller's CUDA graph in pool mode allocate graph-owned memory; after the graph is
    destroyed mfgp_graph_mem_trim returns it (include/mfgp.h).  The calls before, inside and after give the same result."""
    import torch

    from multi_fidelity_gpflow_b200 import _lib

    rng = np.random.default_rng(5)
    X, th = rand_X(rng, 2950, 5), rand_theta(rng, 5)  # 2950 points: the streaming path, which allocates a workspace
    ref = h.cov(X, None, th)
    dev = torch.device("cuda:0")
    tX, tth = torch.from_numpy(X).to(dev), torch.from_numpy(th).to(dev)
    K = torch.empty(len(X), len(X), dtype=torch.float64, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    h.set_async(True)
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            h.set_stream(side.cuda_stream)
            with torch.cuda.graph(g, stream=side, capture_error_mode="relaxed"):
                assert _lib._lib.mfgp_cov(h._h, _lib._ptr(tX), len(X), None, len(X), 5, _lib._ptr(tth), _lib._ptr(K), len(X)) == 0
        h.set_stream(None)
        for _ in range(3):
            K.zero_()
            g.replay()
            torch.cuda.synchronize()
            np.testing.assert_array_equal(K.cpu().numpy(), ref)
        del g
    finally:
        h.set_stream(None)
        h.set_async(False)
    h.graph_mem_trim()
    np.testing.assert_array_equal(h.cov(X, None, th), ref)


def test_workspace_fixed_arena(h):
    """mfgp_workspace: MEASURE records a call's temporaries, FIXED serves them from one arena (eagerly and inside a captured
    graph, which then has no allocation nodes), a call that outgrows the arena spills to the pool; results never change."""
    import torch

    from multi_fidelity_gpflow_b200 import _lib

    rng = np.random.default_rng(6)
    X, Xbig, th = rand_X(rng, 2950, 5), rand_X(rng, 3300, 5), rand_theta(rng, 5)
    ref, ref_big = h.cov(X, None, th), h.cov(Xbig, None, th)
    dev = torch.device("cuda:0")
    tX, tXb, tth = torch.from_numpy(X).to(dev), torch.from_numpy(Xbig).to(dev), torch.from_numpy(th).to(dev)
    K = torch.empty(len(X), len(X), dtype=torch.float64, device=dev)
    Kb = torch.empty(len(Xbig), len(Xbig), dtype=torch.float64, device=dev)
    call = lambda x, k: _lib._lib.mfgp_cov(h._h, _lib._ptr(x), len(x), None, len(x), 5, _lib._ptr(tth), _lib._ptr(k), len(x))
    with pytest.raises(ValueError):
        h.workspace(h.WS_FIXED)  # FIXED follows MEASURE
    try:
        assert h.workspace(h.WS_MEASURE) == 0
        assert call(tX, K) == 0
        need = h.workspace(h.WS_FIXED)
        assert need >= 8 * 14 * 2950  # the prescaled panel workspace P[2d+4][Npad] at least
        for _ in range(2):  # eager calls reuse the arena from its start
            K.zero_()
            assert call(tX, K) == 0 and h.sync() == 0
            np.testing.assert_array_equal(K.cpu().numpy(), ref)
        assert call(tXb, Kb) == 0 and h.sync() == 0  # larger than measured: the excess comes from the pool
        np.testing.assert_array_equal(Kb.cpu().numpy(), ref_big)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        h.set_async(True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            h.set_stream(side.cuda_stream)
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                assert call(tX, K) == 0
        h.set_stream(None)
        for _ in range(3):
            K.zero_()
            g.replay()
            torch.cuda.synchronize()
            np.testing.assert_array_equal(K.cpu().numpy(), ref)
        del g
    finally:
        h.set_stream(None)
        h.set_async(False)
        h.workspace(h.WS_POOL)
    np.testing.assert_array_equal(h.cov(X, None, th), ref)


@pytest.mark.parametrize("M,K,nc", [(1000, 512, 2), (37, 1024, 1), (4096, 1023, 2), (5, 2, 2)])
def test_tall_skinny_update(h, M, K, nc):
    """mfgp_tall_skinny_update (forward-substitution step of the distributed Cholesky) against NumPy."""
    import torch

    from multi_fidelity_gpflow_b200 import _lib

    rng = np.random.default_rng(M + K)
    lda = K + (K & 1)
    A = np.zeros((M, lda))
    A[:, :K] = rng.standard_normal((M, K))
    X, Y = rng.standard_normal((K, 2)), rng.standard_normal((M, 2))
    Ad, Xd, Yd = (torch.from_numpy(a).cuda() for a in (A, X, Y))
    rc = _lib._lib.mfgp_tall_skinny_update(h._h, M, K, nc, -0.7, _lib._ptr(Ad), lda, _lib._ptr(Xd), 2, _lib._ptr(Yd), 2)
    assert rc == 0 and h.sync() == 0
    ref = Y.copy()
    ref[:, :nc] += -0.7 * A[:, :K] @ X[:, :nc]
    np.testing.assert_allclose(Yd.cpu().numpy(), ref, rtol=1e-12, atol=1e-12)
