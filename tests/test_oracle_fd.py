"""Independent checks of the oracle's gradients and predictions (SURVEY 8(c) form 3): the torch-autograd twin
(oracle/mfgp_oracle_torch.py) against CENTRAL FINITE DIFFERENCES of the NumPy forward (oracle/mfgp_oracle.py), and the
closed-form gradient used by the large-N parity cases against both.  These rows have no reference-held known answer
(GPR.predict_f values, HeteroscedasticGaussian, MaskedGaussian, kl_multiplier != 1, minibatch scale != 1, the graph
kernel), so finite differences / a second algebraic route are the only independent authority."""
import numpy as np
import pytest

from oracle import mfgp_oracle as onp
from oracle import mfgp_oracle_torch as otc


def central_fd(f, x, h):
    """Gradient of scalar f at x by central differences with per-coordinate step h * max(1, |x_i|)."""
    x = np.asarray(x, dtype=np.float64)
    g = np.zeros(x.size)
    flat = x.ravel()
    for i in range(x.size):
        step = h * max(1.0, abs(flat[i]))
        xp, xm = flat.copy(), flat.copy()
        xp[i] += step
        xm[i] -= step
        g[i] = (f(xp.reshape(x.shape)) - f(xm.reshape(x.shape))) / (2.0 * step)
    return g.reshape(x.shape)


def directional_fd(f, x, v, h):
    return (f(x + h * v) - f(x - h * v)) / (2.0 * h)


@pytest.mark.parametrize("name", ["forrester", "hbs"])
def test_gpr_gradient_autograd_vs_fd_vs_closed_form(name):
    rng = np.random.default_rng(0)
    if name == "forrester":
        ds = onp.forrester_dataset()
        X, Y, d = ds["X"], ds["Y"], 1
    else:
        ds = onp.load_dataset("hbs")
        X, Y, d = ds["X"], np.ascontiguousarray(ds["Y"][:, :5]), 5
    theta = onp.default_theta(d) * np.exp(0.2 * rng.standard_normal(2 * d + 3))
    noise = 2e-3
    lml, g, gn = otc.gpr_lml_value_and_grad(X, Y, theta, noise)
    assert abs(lml - onp.gpr_lml(X, Y, theta, noise)) < 1e-11 * abs(lml)
    fd = central_fd(lambda th: onp.gpr_lml(X, Y, th, noise), theta, 1e-6)
    np.testing.assert_allclose(g, fd, rtol=2e-6, atol=2e-6 * np.abs(g).max())
    fdn = (onp.gpr_lml(X, Y, theta, noise * (1 + 1e-5)) - onp.gpr_lml(X, Y, theta, noise * (1 - 1e-5))) / (2e-5 * noise)
    assert abs(gn - fdn) < 2e-6 * abs(gn)
    lml2, g2, gn2 = onp.gpr_lml_grad_analytic(X, Y, theta, noise)
    assert abs(lml2 - lml) < 1e-11 * abs(lml)
    np.testing.assert_allclose(g2, g, rtol=1e-9, atol=1e-9 * np.abs(g).max())
    assert abs(gn2 - gn) < 1e-9 * abs(gn)


def test_closed_form_gradient_with_dead_rows_and_more_columns():
    """Quirk Q1 rows (fidelity not in {0, 1}) and P > 1 in the closed form used at large N."""
    ds = onp.synthetic_exact_dataset(300)
    X = ds["X"].copy()
    X[5, -1] = 0.5
    X[17, -1] = np.nan
    rng = np.random.default_rng(1)
    Y = np.hstack([ds["Y"], rng.standard_normal((300, 2))])
    lml, g, gn = otc.gpr_lml_value_and_grad(X, Y, ds["theta"], ds["noise"])
    lml2, g2, gn2 = onp.gpr_lml_grad_analytic(X, Y, ds["theta"], ds["noise"])
    assert abs(lml2 - lml) < 1e-11 * abs(lml)
    np.testing.assert_allclose(g2, g, rtol=1e-9, atol=1e-9 * np.abs(g).max())
    assert abs(gn2 - gn) < 1e-9 * abs(gn)


@pytest.mark.parametrize("variant", ["gaussian_scaled_klmult", "hetero", "masked", "separate"])
def test_svgp_gradient_autograd_vs_fd(variant):
    """Directional derivatives of loss = -ELBO + (kl_mult - 1) KL along random directions of every parameter group."""
    ds = onp.load_dataset("hbs")
    rng = np.random.default_rng(3)
    idx = rng.permutation(53)[:20]
    X, Y = ds["X"][idx], np.ascontiguousarray(ds["Y"][idx][:, :6])
    M, d, P = 12, 5, 6
    Z = ds["Z_kmeans50"][:M].copy()
    mixing = variant != "separate"
    L = 3 if mixing else P
    W = 0.3 * rng.standard_normal((P, L)) if mixing else None
    thetas = np.tile(onp.default_theta(d), (L, 1)) * np.exp(0.2 * rng.standard_normal((L, 2 * d + 3)))
    q_mu = 0.3 * rng.standard_normal((M, L))
    q_sqrt = np.tile(0.3 * np.eye(M), (L, 1, 1)) + 0.05 * np.tril(rng.standard_normal((L, M, M)))
    hetero, masked, kl_mult, num_data = False, False, 1.0, None
    lik = 0.7
    if variant == "gaussian_scaled_klmult":
        kl_mult, num_data = 2.5, 53
    elif variant == "hetero":
        hetero, num_data = True, 53
        Y = np.hstack([Y, 0.1 + 0.2 * rng.random(Y.shape)])
    elif variant == "masked":
        masked, num_data = True, 53
        Y = Y.copy()
        Y[rng.random(Y.shape) < 0.3] = np.nan
        lik = 0.5 + rng.random(P)

    def loss(Z_=Z, th_=thetas, qm_=q_mu, qs_=q_sqrt, lv_=lik, W_=W):
        e, k = onp.svgp_elbo(X, Y, Z_, th_, qm_, qs_, lv_, W_, num_data, hetero, masked)
        return -e + (kl_mult - 1.0) * k

    r = otc.svgp_value_and_grad(X, Y, Z, thetas, q_mu, q_sqrt, lik, W, num_data, hetero, kl_mult, masked)
    assert abs(r["loss"] - loss()) < 1e-11 * abs(r["loss"])
    groups = [("g_Z", Z, "Z_"), ("g_thetas", thetas, "th_"), ("g_q_mu", q_mu, "qm_"), ("g_q_sqrt", q_sqrt, "qs_"),
              ("g_lik_var", np.asarray(lik, dtype=np.float64), "lv_")]
    if mixing:
        groups.append(("g_W", W, "W_"))
    for key, x0, kw in groups:
        for _ in range(2):
            v = rng.standard_normal(np.shape(x0))
            if key == "g_q_sqrt":
                v = np.tril(v)
            if key == "g_Z":
                v[:, -1] = 0.0  # the fidelity column is not a continuous variable (quirk Q5)
            want = float(np.sum(np.asarray(r[key]) * v))
            got = directional_fd(lambda x: loss(**{kw: x}), x0, v, 1e-6)
            assert abs(got - want) < 5e-6 * max(abs(want), 1e-3 * np.abs(np.asarray(r[key])).max() * np.abs(v).sum()), (key, got, want)


def test_gpr_predict_against_a_second_algebraic_route():
    """GPR.predict_f has no reference-held value: check the Cholesky-based restatement against plain linear solves."""
    ds = onp.forrester_dataset()
    X, Y = ds["X"], ds["Y"]
    theta, noise = onp.default_theta(1) * np.array([1.7, 0.3, 1.4, 0.5, 0.8]), 1e-3
    for Xs in (ds["X_plot_L"][::7], ds["X_plot_H"][::7]):
        mean, var = onp.gpr_predict(X, Y, Xs, theta, noise)
        Kn = onp.mf_K(X, None, theta) + noise * np.eye(X.shape[0])
        Ks = onp.mf_K(X, Xs, theta)
        sol = np.linalg.solve(Kn, np.hstack([Y, Ks]))
        np.testing.assert_allclose(mean, Ks.T @ sol[:, :1], rtol=1e-9, atol=1e-9 * np.abs(mean).max())
        np.testing.assert_allclose(var, onp.mf_K_diag(Xs, theta) - np.sum(Ks * sol[:, 1:], axis=0), rtol=1e-7, atol=1e-9)
