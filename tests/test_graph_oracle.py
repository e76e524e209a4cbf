"""CPU checks of the graph-kernel oracle (reference mfgpflow/graph.py): the NumPy block-by-block restatement against the
masked torch form, structural properties the reference code implies, and the gradient convention (lower-triangle Cholesky
+ symmetrised sensitivity, as TensorFlow's registered Cholesky gradient) against finite differences where it is a true
gradient (symmetric K)."""
import numpy as np

from oracle import mfgp_oracle as onp
from oracle import mfgp_oracle_torch as otc


def graph_problem(rng, N=40, d=3, m=2, P=2, symmetric=False, noise=1e-3, max_draws=200):
    """Random graph-model problem.  The reference's construction is not positive semi-definite in general (K_HH has no
    rho_i rho_j rho_LF_ij cross terms, graph.py:84): draw short length-scales / weak cross-correlations and redraw (a bounded
    number of times) until the matrix the Cholesky sees -- lower triangle of K + noise I -- is positive definite with margin."""
    for _ in range(max_draws):
        X = np.hstack([rng.random((N, d)), rng.integers(0, m + 1, size=(N, 1)).astype(float)])
        Y = rng.standard_normal((N, P))
        rho = 0.5 + rng.random(m)
        rho_LF = 0.02 + 0.08 * rng.random((m, m))
        kL = [(0.3 + 0.2 * rng.random(d), 0.8 + 0.4 * rng.random()) for _ in range(m)]
        if symmetric:  # identical LF kernels and a symmetric rho_LF: K symmetric, so the TF convention is the true gradient
            rho_LF = 0.5 * (rho_LF + rho_LF.T)
            kL = [kL[0]] * m
        kD = (0.3 + 0.2 * rng.random(d), 0.5 + rng.random())
        gth = onp.graph_pack(rho, rho_LF, kL, kD)
        K = onp.graph_K(X, gth, m)
        if np.linalg.eigvalsh(np.tril(K) + np.tril(K, -1).T).min() + noise > 0.3 * noise:
            return X, Y, gth
    raise RuntimeError(f"no positive-definite graph-kernel problem in {max_draws} draws (N={N}, d={d}, m={m})")


def test_numpy_and_torch_forms_agree_and_structure():
    rng = np.random.default_rng(0)
    for m in (1, 2, 3):
        X, Y, gth = graph_problem(rng, m=m)
        X[3, -1] = 7.0      # not a fidelity level: dead row (graph.py:54 zeros)
        X[9, -1] = np.nan
        K = onp.graph_K(X, gth, m)
        Kt = otc.graph_K(X, otc._t(gth), m).numpy()
        np.testing.assert_allclose(K, Kt, rtol=1e-13, atol=1e-15)
        dead = [3, 9]
        off = K - np.diag(np.diag(K))
        assert np.all(off[dead] == 0) and np.all(off[:, dead] == 0) and np.allclose(np.diag(K)[dead], 1e-6)
        np.testing.assert_allclose(np.diag(K) - 1e-6, onp.graph_K_diag(X, gth, m), rtol=1e-13)
        if m > 1:
            assert not np.allclose(K, K.T)  # the LF-LF cross blocks use the ROW source's kernel (graph.py:63)
        # one LF source, rho_LF unused: the graph kernel is the linear multi-fidelity kernel + jitter (linear.py:55-104)
    X, Y, gth = graph_problem(rng, m=1)
    d = X.shape[1] - 1
    rho, _, kL, kD = onp.graph_unpack(gth, 1, d)
    th = onp.pack_theta(rho[0], kL[0][0], kL[0][1], kD[0], kD[1])
    np.testing.assert_allclose(onp.graph_K(X, gth, 1), onp.mf_K(X, None, th) + 1e-6 * np.eye(X.shape[0]), rtol=1e-13, atol=1e-15)
    v = onp.graph_gpr_lml(X, Y, gth, 1, 1e-3)
    assert abs(v - onp.gpr_lml(X, Y, th, 1e-3 + 1e-6)) < 1e-11 * abs(v)


def test_gradient_convention():
    rng = np.random.default_rng(1)
    m, d = 2, 3
    # value: torch (asymmetric K straight into the Cholesky) == NumPy (explicit lower-triangle symmetrisation)
    X, Y, gth = graph_problem(rng, m=m, d=d)
    v, g, gn = otc.graph_gpr_lml_value_and_grad(X, Y, gth, m, 1e-3)
    assert abs(v - onp.graph_gpr_lml(X, Y, gth, m, 1e-3)) < 1e-11 * abs(v)
    assert g[m] == 0 and g[m + m * m - 1] == 0  # rho_LF diagonal is never read
    assert g[m + 1] != g[m + 2]                   # rho_LF[0, 1] and rho_LF[1, 0] scale different kernels
    # symmetric configuration: the convention is the true gradient -> central finite differences of the NumPy value
    X, Y, gth = graph_problem(rng, m=m, d=d, symmetric=True)
    v, g, gn = otc.graph_gpr_lml_value_and_grad(X, Y, gth, m, 1e-3)
    f = lambda t: onp.graph_gpr_lml(X, Y, t, m, 1e-3)
    o = m + m * m
    groups = [[q] for q in range(m)]                                         # rho_i
    groups += [[o + q, o + (d + 1) + q] for q in range(d + 1)]                # the tied LF kernels move together (K stays symmetric)
    groups += [[o + 2 * (d + 1) + q] for q in range(d + 1)]                   # delta kernel
    for grp in groups:
        e = np.zeros_like(gth)
        e[grp] = 1e-6 * max(1.0, abs(gth[grp[0]]))
        fd = (f(gth + e) - f(gth - e)) / (2 * e[grp[0]])
        want = g[grp].sum()
        assert abs(fd - want) < 5e-6 * max(abs(want), 1e-3 * np.abs(g).max()), (grp, fd, want)
    # moving rho_LF[0,1] and rho_LF[1,0] together keeps K symmetric: FD of the pair = sum of the two entries
    e = np.zeros_like(gth)
    e[m + 1] = e[m + 2] = 1e-6
    fd = (f(gth + e) - f(gth - e)) / 2e-6
    assert abs(fd - (g[m + 1] + g[m + 2])) < 5e-6 * abs(fd)
