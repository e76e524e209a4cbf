"""GPU parity tests of the SVGP ELBO / gradient / predict path (C-ABI) against the torch oracle."""
import numpy as np
import pytest

from oracle import mfgp_oracle as onp
from oracle import mfgp_oracle_torch as otc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h():
    from multi_fidelity_gpflow_b200 import _lib

    return _lib.Handle(0)


def make_problem(rng, X, Y, Z, L, d, mixing, hetero=False, perturb=True):
    M, P = Z.shape[0], Y.shape[1]
    th = np.tile(onp.default_theta(d), (L, 1))
    q_mu = np.zeros((M, L))
    q_sqrt = np.tile(0.1 * np.eye(M), (L, 1, 1))
    W = onp.initialize_W(P, L, 0.4, 0.2) if mixing else None
    Zp = Z.copy()
    if perturb:
        th = th * np.exp(0.2 * rng.standard_normal(th.shape))
        q_mu = 0.3 * rng.standard_normal((M, L))
        q_sqrt = q_sqrt + 0.02 * np.tril(rng.standard_normal((L, M, M)))
        Zp[:, :-1] += 0.01 * rng.standard_normal((M, d))
        if mixing:
            W = W + 0.05 * rng.standard_normal(W.shape)
    Yh = np.hstack([Y, 0.1 + 0.2 * rng.random(Y.shape)]) if hetero else Y
    return dict(X=X, Y=Yh, Z=Zp, thetas=th, q_mu=q_mu, q_sqrt=q_sqrt, W=W)


def compare(h, pr, lik_var, num_data=None, kl_mult=1.0, hetero=False, vtol=1e-9, gtol=1e-7, masked=False):
    B = pr["X"].shape[0]
    scale = 1.0 if num_data is None else num_data / B
    got = h.svgp_elbo_grad(pr["X"], pr["Y"], pr["Z"], pr["thetas"], pr["W"], pr["q_mu"], pr["q_sqrt"], lik_var,
                           scale=scale, kl_mult=kl_mult, hetero=hetero, masked=masked)
    ref = otc.svgp_value_and_grad(pr["X"], pr["Y"], pr["Z"], pr["thetas"], pr["q_mu"], pr["q_sqrt"], lik_var, pr["W"],
                                  num_data, hetero, kl_mult, masked)
    assert abs(got["elbo"] - ref["elbo"]) <= vtol * abs(ref["elbo"]), (got["elbo"], ref["elbo"])
    assert abs(got["kl"] - ref["kl"]) <= vtol * max(1.0, abs(ref["kl"]))
    for k in ("g_q_mu", "g_q_sqrt", "g_Z", "g_thetas", "g_W"):
        if ref[k] is None:
            continue
        scale_k = np.abs(ref[k]).max()
        np.testing.assert_allclose(got[k], ref[k], rtol=gtol, atol=gtol * max(scale_k, 1e-30), err_msg=k)
    np.testing.assert_allclose(got["g_lik_var"], ref["g_lik_var"], rtol=gtol, atol=gtol * np.abs(ref["g_lik_var"]).max())
    assert np.all(got["g_Z"][:, -1] == 0)  # quirk Q5: fidelity column of Z gets exactly zero gradient
    return got, ref


def test_hbs_singlebin_init_matches_golden_setup(h):
    """C3 config at init: ELBO recorded in SURVEY App. B note on G4."""
    ds = onp.load_dataset("hbs")
    pr = make_problem(np.random.default_rng(0), ds["X"], ds["Y"], ds["Z_kmeans50"], 49, 5, False, perturb=False)
    got, _ = compare(h, pr, 1.0)
    assert abs(got["elbo"] + 7351.274738200964) < 1e-9 * 7351.27


def test_hbs_singlebin_perturbed(h):
    ds = onp.load_dataset("hbs")
    pr = make_problem(np.random.default_rng(1), ds["X"], ds["Y"], ds["Z_kmeans50"], 49, 5, False)
    compare(h, pr, 0.7)


@pytest.mark.parametrize("kl_mult,num_data", [(1.0, 53), (2.5, 53), (1.0, 530)])
def test_hbs_latent_mixing(h, kl_mult, num_data):
    ds = onp.load_dataset("hbs")
    pr = make_problem(np.random.default_rng(2), ds["X"], ds["Y"], ds["Z_kmeans50"], 10, 5, True)
    compare(h, pr, 0.9, num_data=num_data, kl_mult=kl_mult)


def test_hbs_latent_heteroscedastic(h):
    ds = onp.load_dataset("hbs")
    pr = make_problem(np.random.default_rng(3), ds["X"], ds["Y"], ds["Z_kmeans50"], 10, 5, True, hetero=True)
    compare(h, pr, 0.5, num_data=53, hetero=True)


@pytest.mark.parametrize("mixing,L", [(True, 10), (False, 49)])
def test_masked_gaussian_missing_outputs(h, mixing, L):
    """MaskedGaussian (reference notebooks/"demo: missing output.ipynb" cell 2) -- NaN entries of Y are
    missing outputs, one likelihood variance per output.  ~30 % of the entries missing, one output missing entirely."""
    ds = onp.load_dataset("hbs")
    rng = np.random.default_rng(11)
    pr = make_problem(rng, ds["X"], ds["Y"], ds["Z_kmeans50"], L, 5, mixing)
    Y = pr["Y"].copy()
    Y[rng.random(Y.shape) < 0.3] = np.nan
    Y[:, 7] = np.nan
    pr["Y"] = Y
    lik = 0.4 + rng.random(49)
    got, ref = compare(h, pr, lik, num_data=53, masked=True)
    assert got["g_lik_var"].shape == (49,) and got["g_lik_var"][7] == 0.0  # a never-observed output has no likelihood gradient
    # no NaN in Y: the masked epilogue with equal variances is the plain Gaussian one
    pr["Y"] = ds["Y"]
    a = h.svgp_elbo_grad(pr["X"], pr["Y"], pr["Z"], pr["thetas"], pr["W"], pr["q_mu"], pr["q_sqrt"], np.full(49, 0.8), masked=True)
    b = h.svgp_elbo_grad(pr["X"], pr["Y"], pr["Z"], pr["thetas"], pr["W"], pr["q_mu"], pr["q_sqrt"], 0.8)
    assert a["elbo"] == b["elbo"]
    np.testing.assert_allclose(a["g_lik_var"].sum(), b["g_lik_var"], rtol=1e-12)
    np.testing.assert_array_equal(a["g_thetas"], b["g_thetas"])


def test_minibatch_rows(h):
    ds = onp.load_dataset("hbs")
    rng = np.random.default_rng(4)
    idx = rng.permutation(53)[:17]
    pr = make_problem(rng, ds["X"][idx], ds["Y"][idx], ds["Z_kmeans50"], 10, 5, True)
    compare(h, pr, 1.1, num_data=53)


def test_goku_latent_full(h):
    """C4: L=15, M=300, B=1164, P=64 (M > 128 exercises the blocked potrf / trtri merge levels)."""
    ds = onp.load_dataset("goku")
    pr = make_problem(np.random.default_rng(5), ds["X"], ds["Y"], ds["Z_kmeans300"], 15, 10, True)
    compare(h, pr, 1.0, num_data=1164)


def test_goku_singlebin_subset_with_dead_inducing_points(h):
    """8 of the 64 bins; Z contains the two 1-ulp-off fidelity centres (quirk Q1)."""
    ds = onp.load_dataset("goku")
    pr = make_problem(np.random.default_rng(6), ds["X"], ds["Y"][:, :8], ds["Z_kmeans300"], 8, 10, False, perturb=False)
    compare(h, pr, 1.0)


def test_svgp_predict(h):
    ds = onp.load_dataset("hbs")
    rng = np.random.default_rng(7)
    for mixing, L in ((False, 49), (True, 10)):
        pr = make_problem(rng, ds["X"], ds["Y"], ds["Z_kmeans50"], L, 5, mixing)
        mean, var = h.svgp_predict(ds["X_test"], pr["Z"], pr["thetas"], pr["W"], pr["q_mu"], pr["q_sqrt"])
        rm, rv = onp.svgp_predict(ds["X_test"], pr["Z"], pr["thetas"], pr["q_mu"], pr["q_sqrt"], pr["W"])
        assert mean.shape == (10, 49) and var.shape == (10, 49)  # tests/test_ho2021_singlebin.py:87-88
        np.testing.assert_allclose(mean, rm, rtol=1e-8, atol=1e-8 * np.abs(rm).max())
        np.testing.assert_allclose(var, rv, rtol=1e-8, atol=1e-8 * np.abs(rv).max())
