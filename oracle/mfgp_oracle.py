"""CPU oracle for the Kennedy-O'Hagan linear multi-fidelity GP hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``multi_fidelity_gpflow_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg do.  The product path is the CUDA library and fails loudly
without it.

What is restated here (NumPy/SciPy float64, no autograd -- gradients live in
``mfgp_oracle_torch.py``):

* the reference's own kernel, ``mfgpflow/linear.py:55-136`` (gather by fidelity mask,
  five SquaredExponential blocks, scatter into a zero matrix);
* the GPflow 2.9.0 code the reference calls (third-party, pinned in
  ``requirements.txt:2``; not vendored under /root/reference, so its published
  algorithm is restated): ``GPR.log_marginal_likelihood`` / ``predict_f``
  (call sites ``linear.py:206,227``), ``SVGP.elbo`` / ``prior_kl`` / ``predict_f``
  (call sites ``singlebin_svgp.py:83,97``, ``linear_svgp.py:177,184,188``),
  ``conditionals.util.base_conditional`` / ``mix_latent_gp``, ``gauss_kl``,
  ``Gaussian._variational_expectations`` and the reference's
  ``HeteroscedasticGaussian`` (``linear_svgp.py:243-267``);
* Keras/TF-2.10 ``Adam`` + ``CosineDecay`` exactly as the reference's ``optimize()``
  loops drive them (``linear.py:201-209``, ``singlebin_svgp.py:77-85``), including the
  float32-rounded hyper-parameters.

Parity pin: the values recorded in the reference's notebooks (G1-G7, see
``tests/golden/goldens.json`` and ``tests/test_oracle_goldens.py``) are reproduced by
this file to <= 1e-12 relative.  ``GPR.predict_f`` values, the heteroscedastic
likelihood and ``kl_multiplier != 1`` are NOT pinned by any reference artefact
("parity unpinned by reference; pinned by oracle self-consistency + finite
differences").
"""
from __future__ import annotations

import math
import os

import numpy as np
import scipy.linalg as sla

JITTER = 1e-6  # gpflow.config.default_jitter()
LIK_VAR_LOWER = 1e-6  # gpflow.likelihoods.Gaussian.DEFAULT_VARIANCE_LOWER_BOUND
LOG2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# transforms (gpflow.utilities.positive == softplus, lower bound 0)
# --------------------------------------------------------------------------------------
def softplus(u):
    return np.logaddexp(0.0, u)


def softplus_inv(x):
    x = np.asarray(x, dtype=np.float64)
    return x + np.log(-np.expm1(-x))


def fill_triangular_inverse(L):
    """tfp.math.fill_triangular_inverse for one lower-triangular [M, M] matrix.

    Only used to map q_sqrt <-> its unconstrained vector; the ordering does not affect
    Adam (element-wise), so any fixed bijection tril -> R^{M(M+1)/2} would do.
    """
    M = L.shape[-1]
    idx = np.tril_indices(M)
    return L[..., idx[0], idx[1]]


# --------------------------------------------------------------------------------------
# theta packing: [rho, ls_L[d], var_L, ls_delta[d], var_delta]   (length 2d+3)
# --------------------------------------------------------------------------------------
def pack_theta(rho, ls_L, var_L, ls_d, var_d):
    return np.concatenate([[rho], np.atleast_1d(ls_L), [var_L], np.atleast_1d(ls_d), [var_d]]).astype(np.float64)


def unpack_theta(theta, d):
    theta = np.asarray(theta, dtype=np.float64)
    assert theta.shape[-1] == 2 * d + 3
    return theta[0], theta[1 : 1 + d], theta[1 + d], theta[2 + d : 2 + 2 * d], theta[2 + 2 * d]


def default_theta(d):
    return pack_theta(1.0, np.ones(d), 1.0, np.ones(d), 1.0)


# --------------------------------------------------------------------------------------
# gpflow.kernels.SquaredExponential (stationaries.py / utilities/ops.py::square_distance)
# --------------------------------------------------------------------------------------
def square_distance(A, B):
    As = np.sum(A * A, axis=-1)
    Bs = np.sum(B * B, axis=-1)
    dist = -2.0 * (A @ B.T)
    dist += As[:, None] + Bs[None, :]
    return dist


def se_K(A, B, ls, var):
    return var * np.exp(-0.5 * square_distance(A / ls, B / ls))


# --------------------------------------------------------------------------------------
# reference kernel, mfgpflow/linear.py:55-136
# --------------------------------------------------------------------------------------
def mf_K(X, X2, theta):
    """LinearMultiFidelityKernel.K  (linear.py:55-104), rho = rho[0, 0]."""
    X = np.asarray(X, dtype=np.float64)
    X2 = X if X2 is None else np.asarray(X2, dtype=np.float64)
    d = X.shape[1] - 1
    rho, lsL, vL, lsD, vD = unpack_theta(theta, d)
    mL, mH = np.where(X[:, -1] == 0)[0], np.where(X[:, -1] == 1)[0]  # :67-68
    m2L, m2H = np.where(X2[:, -1] == 0)[0], np.where(X2[:, -1] == 1)[0]  # :69-70
    XL, XH, X2L, X2H = X[mL, :-1], X[mH, :-1], X2[m2L, :-1], X2[m2H, :-1]  # :73-76
    K = np.zeros((X.shape[0], X2.shape[0]))  # :82
    K[np.ix_(mL, m2L)] = se_K(XL, X2L, lsL, vL)  # :93, :99
    K[np.ix_(mL, m2H)] = se_K(XL, X2H, lsL, vL) * rho  # :94, :100
    K[np.ix_(mH, m2L)] = se_K(XH, X2L, lsL, vL) * rho  # :95, :101
    K[np.ix_(mH, m2H)] = se_K(XH, X2H, lsL, vL) * (rho * rho) + se_K(XH, X2H, lsD, vD)  # :96, :102
    return K


def mf_K_diag(X, theta):
    """LinearMultiFidelityKernel.K_diag (linear.py:106-136)."""
    X = np.asarray(X, dtype=np.float64)
    d = X.shape[1] - 1
    rho, _, vL, _, vD = unpack_theta(theta, d)
    out = np.zeros(X.shape[0])
    out[X[:, -1] == 0] = vL
    out[X[:, -1] == 1] = vL * (rho * rho) + vD  # :127  K_diag_L * rho**2 + K_diag_delta
    return out


# --------------------------------------------------------------------------------------
# reference graph kernel, mfgpflow/graph.py:39-115 (several LF sources)
# --------------------------------------------------------------------------------------
def graph_unpack(gtheta, m, d):
    """gtheta layout of include/mfgp.h: rho [m], rho_LF [m, m], (ls_Li [d], var_Li) per source, ls_delta [d], var_delta."""
    g = np.asarray(gtheta, dtype=np.float64)
    rho, rho_LF = g[:m], g[m:m + m * m].reshape(m, m)
    o = m + m * m
    ks = []
    for _ in range(m + 1):
        ks.append((g[o:o + d], g[o + d]))
        o += d + 1
    return rho, rho_LF, ks[:m], ks[m]


def graph_pack(rho, rho_LF, kernels_L, kernel_delta):
    parts = [np.ravel(rho), np.ravel(rho_LF)]
    for ls, var in list(kernels_L) + [kernel_delta]:
        parts += [np.ravel(ls), [var]]
    return np.concatenate([np.asarray(p, dtype=np.float64) for p in parts])


def graph_K(X, gtheta, m):
    """GraphMultiFidelityKernel.K(X) for X2 = X (graph.py:39-93), block by block in the reference's order, rho = rho[:, 0]."""
    X = np.asarray(X, dtype=np.float64)
    N, d = X.shape[0], X.shape[1] - 1
    rho, rho_LF, kL, kD = graph_unpack(gtheta, m, d)
    masks_L = [np.where(X[:, -1] == i)[0] for i in range(m)]  # :45
    mask_H = np.where(X[:, -1] == m)[0]                       # :46
    K = np.zeros((N, N))                                      # :54
    for i in range(m):                                        # :57-66  LF-LF, the ROW source's kernel
        for j in range(m):
            Xi, Xj = X[masks_L[i], :-1], X[masks_L[j], :-1]
            rho_ij = rho_LF[i, j] if i != j else 1.0
            K[np.ix_(masks_L[i], masks_L[j])] = rho_ij * se_K(Xi, Xj, *kL[i])
    if mask_H.size > 0:                                       # :69, :82
        XH = X[mask_H, :-1]
        for i in range(m):                                    # :70-79  LF-HF and HF-LF
            XL = X[masks_L[i], :-1]
            K[np.ix_(masks_L[i], mask_H)] = se_K(XL, XH, *kL[i]) * rho[i]
            K[np.ix_(mask_H, masks_L[i])] = se_K(XH, XL, *kL[i]) * rho[i]
        KHH = sum(se_K(XH, XH, *kL[i]) * rho[i] ** 2 for i in range(m)) + se_K(XH, XH, *kD)  # :84-85
        K[np.ix_(mask_H, mask_H)] = KHH
    return K + np.eye(N) * 1e-6                               # :91


def graph_K_diag(X, gtheta, m):
    """GraphMultiFidelityKernel.K_diag (graph.py:96-115)."""
    X = np.asarray(X, dtype=np.float64)
    d = X.shape[1] - 1
    rho, _, kL, kD = graph_unpack(gtheta, m, d)
    out = np.zeros(X.shape[0])
    for i in range(m):
        out[X[:, -1] == i] = kL[i][1]
    out[X[:, -1] == m] = sum(kL[i][1] * rho[i] ** 2 for i in range(m)) + kD[1]
    return out


def graph_gpr_lml(X, Y, gtheta, m, noise):
    """GPR.log_marginal_likelihood with the graph kernel (model graph.py:118-141).  tf.linalg.cholesky reads the LOWER
    triangle of K + noise I only (K need not be symmetric, graph.py:63); so does np.linalg.cholesky."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    N = X.shape[0]
    L = np.linalg.cholesky(np.tril(graph_K(X, gtheta, m) + noise * np.eye(N)) + np.tril(graph_K(X, gtheta, m), -1).T)
    A = sla.solve_triangular(L, Y, lower=True)
    return float(np.sum(-0.5 * np.sum(A * A, axis=0) - 0.5 * N * LOG2PI - np.sum(np.log(np.diag(L)))))


# --------------------------------------------------------------------------------------
# GPflow GPR (models/gpr.py, logdensities.py::multivariate_normal)
# --------------------------------------------------------------------------------------
def gpr_lml(X, Y, theta, noise):
    """GPR.log_marginal_likelihood: one K / one Cholesky shared by all P columns."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    N = X.shape[0]
    Kn = mf_K(X, None, theta) + noise * np.eye(N)
    L = np.linalg.cholesky(Kn)
    A = sla.solve_triangular(L, Y, lower=True)
    p = -0.5 * np.sum(A * A, axis=0)
    p -= 0.5 * N * LOG2PI
    p -= np.sum(np.log(np.diag(L)))
    return float(np.sum(p))


def gpr_lml_grad_analytic(X, Y, theta, noise):
    """LML and its gradient w.r.t. the CONSTRAINED [theta (2d+3), noise] in closed form:
    G = alpha alpha^T - P K_n^-1, dLML/dtheta = 1/2 sum_ij G_ij dK_ij/dtheta, dLML/dnoise = 1/2 tr G, with the kernel
    derivatives of App. A.1 (mfgpflow/linear.py:93-102 differentiated by hand).  It is what tape.gradient (linear.py:207)
    returns, without autograd's memory: the large-N parity cases (N > 12 288) use it; tests/test_oracle_fd.py checks it
    against the torch-autograd twin and central finite differences at small N.  Returns (lml, g_theta, g_noise)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    N, d, P = X.shape[0], X.shape[1] - 1, Y.shape[1]
    rho, lsL, vL, lsD, vD = unpack_theta(theta, d)
    f = X[:, -1]
    s = np.where(f == 0, 1.0, np.where(f == 1, rho, 0.0))
    hidx = np.where(f == 1)[0]
    x = np.where(((f == 0) | (f == 1))[:, None], X[:, :-1], 0.0)
    KL = se_K(x, x, lsL, vL)
    K = KL * s[:, None]
    K *= s[None, :]
    xh = x[hidx]
    KD = se_K(xh, xh, lsD, vD)
    K[np.ix_(hidx, hidx)] += KD
    K[np.diag_indices(N)] += noise
    c = sla.cho_factor(K, lower=True, overwrite_a=True, check_finite=False)
    alpha = sla.cho_solve(c, Y, check_finite=False)
    lml = float(-0.5 * np.sum(Y * alpha) - P * (0.5 * N * LOG2PI + np.sum(np.log(np.diag(c[0])))))
    G = sla.cho_solve(c, np.eye(N), check_finite=False)  # K_n^-1
    del c, K
    G *= -float(P)
    G += alpha @ alpha.T
    g = np.zeros(2 * d + 3)
    g_noise = 0.5 * float(np.trace(G))
    GD = G[np.ix_(hidx, hidx)] * KD  # discrepancy terms live on the HF x HF block only
    KL *= G  # from here on KL holds G o K_L
    # d(s_i s_j)/d rho = h_i s_j + s_i h_j
    hs = np.zeros(N)
    hs[hidx] = 1.0
    g[0] = 0.5 * (2.0 * float(hs @ (KL @ s)))
    KL *= s[:, None]
    KL *= s[None, :]  # T^L = G o (s s^T) o K_L
    g[1 + d] = 0.5 * float(KL.sum()) / vL
    g[2 + 2 * d] = 0.5 * float(GD.sum()) / vD
    rs, rsD = KL.sum(axis=1), GD.sum(axis=1)
    for k in range(d):
        # sum_ij T_ij (x_i - x_j)^2 = 2 sum_i rowsum_i x_i^2 - 2 x^T T x   (T symmetric)
        xk, xhk = x[:, k], xh[:, k]
        g[1 + k] = 0.5 * (2.0 * float(rs @ (xk * xk)) - 2.0 * float(xk @ (KL @ xk))) / lsL[k] ** 3
        g[2 + d + k] = 0.5 * (2.0 * float(rsD @ (xhk * xhk)) - 2.0 * float(xhk @ (GD @ xhk))) / lsD[k] ** 3
    return lml, g, g_noise


def gpr_predict(X, Y, Xnew, theta, noise):
    """GPR.predict_f(full_cov=False): base_conditional(white=False).  var is [N*] (same for every column)."""
    X = np.asarray(X, dtype=np.float64)
    N = X.shape[0]
    kmm = mf_K(X, None, theta) + noise * np.eye(N)
    kmn = mf_K(X, Xnew, theta)
    knn = mf_K_diag(Xnew, theta)
    Lm = np.linalg.cholesky(kmm)
    A = sla.solve_triangular(Lm, kmn, lower=True)
    fvar = knn - np.sum(A * A, axis=0)
    A = sla.solve_triangular(Lm.T, A, lower=False)
    fmean = A.T @ np.asarray(Y, dtype=np.float64)
    return fmean, fvar


def gpr_batched_lml(X, Y, thetas, noises):
    """North-star "one GP per k-bin": bin b uses theta[b], noise[b], y = Y[:, b]."""
    return np.array([gpr_lml(X, Y[:, b : b + 1], thetas[b], noises[b]) for b in range(Y.shape[1])])


# --------------------------------------------------------------------------------------
# GPflow SVGP (whiten=True), SeparateIndependent / LinearCoregionalization with
# SharedIndependentInducingVariables
# --------------------------------------------------------------------------------------
def prior_kl(q_mu, q_sqrt):
    """gauss_kl(q_mu, q_sqrt, K=None): q_mu [M, L], q_sqrt [L, M, M]."""
    M, L = q_mu.shape
    Lq = np.tril(q_sqrt)
    diag = np.diagonal(Lq, axis1=-2, axis2=-1)
    two_kl = np.sum(q_mu * q_mu) - M * L - np.sum(np.log(diag * diag)) + np.sum(Lq * Lq)
    return 0.5 * float(two_kl)


def svgp_latent_conditional(Xb, Z, thetas, q_mu, q_sqrt):
    """Per-latent base_conditional(white=True): returns g_mean [B, L], g_var [B, L]."""
    Lnum = thetas.shape[0]
    M = Z.shape[0]
    B = Xb.shape[0]
    g_mean = np.empty((B, Lnum))
    g_var = np.empty((B, Lnum))
    for l in range(Lnum):
        Kmm = mf_K(Z, None, thetas[l]) + JITTER * np.eye(M)
        Kmn = mf_K(Z, Xb, thetas[l])
        knn = mf_K_diag(Xb, thetas[l])
        Lm = np.linalg.cholesky(Kmm)
        A = sla.solve_triangular(Lm, Kmn, lower=True)
        fvar = knn - np.sum(A * A, axis=0)
        Lq = np.tril(q_sqrt[l])
        LTA = Lq.T @ A
        fvar = fvar + np.sum(LTA * LTA, axis=0)
        g_mean[:, l] = A.T @ q_mu[:, l]
        g_var[:, l] = fvar
    return g_mean, g_var


def svgp_predict(Xb, Z, thetas, q_mu, q_sqrt, W=None):
    """SVGP.predict_f(full_cov=False, full_output_cov=False)."""
    g_mean, g_var = svgp_latent_conditional(Xb, Z, thetas, q_mu, q_sqrt)
    if W is None:
        return g_mean, g_var
    return g_mean @ W.T, g_var @ (W * W).T  # conditionals/util.py::mix_latent_gp


def gaussian_var_exp(Y, f_mean, f_var, lik_var, hetero=False):
    """Gaussian._variational_expectations summed over outputs -> [B].

    hetero: Y = [Y_obs | Y_unc], effective variance = lik_var + Y_unc**2 (linear_svgp.py:259).
    """
    P = f_mean.shape[1]
    if hetero:
        Yo, Yu = Y[:, :P], Y[:, P:]
        assert Yu.shape[1] == P
        ev = lik_var + Yu * Yu
    else:
        Yo = Y
        ev = np.broadcast_to(np.asarray(lik_var, dtype=np.float64), f_mean.shape)
    ve = -0.5 * LOG2PI - 0.5 * np.log(ev) - 0.5 * ((Yo - f_mean) ** 2 + f_var) / ev
    return np.sum(ve, axis=-1)


def masked_gaussian_var_exp(Y, f_mean, f_var, lik_var):
    """MaskedGaussian._variational_expectations (reference notebooks/"demo: missing output.ipynb", cell 2): NaN entries of Y
    are missing outputs.  Follows the cell line by line: mask = ~isnan(Y); Y, Fmu, Fvar filled with 0 where masked; the
    closed-form Gaussian VE with the (per-output, shape [P]) variance; masked entries set to 0; sum over outputs -> [B]."""
    mask = ~np.isnan(Y)
    Yf = np.where(mask, Y, 0.0)
    Fm = np.where(mask, f_mean, 0.0)
    Fv = np.where(mask, f_var, 0.0)
    var = np.asarray(lik_var, dtype=np.float64)
    ve = -0.5 * np.log(2.0 * np.pi) - 0.5 * np.log(var) - 0.5 * ((Yf - Fm) ** 2 + Fv) / var
    return np.sum(np.where(mask, ve, 0.0), axis=-1)


def svgp_elbo(Xb, Yb, Z, thetas, q_mu, q_sqrt, lik_var, W=None, num_data=None, hetero=False, masked=False):
    """SVGP.elbo((X, Y)).  Returns (elbo, kl)."""
    kl = prior_kl(q_mu, q_sqrt)
    f_mean, f_var = svgp_predict(Xb, Z, thetas, q_mu, q_sqrt, W)
    ve = masked_gaussian_var_exp(Yb, f_mean, f_var, lik_var) if masked else gaussian_var_exp(Yb, f_mean, f_var, lik_var, hetero)
    scale = 1.0 if num_data is None else float(num_data) / Xb.shape[0]
    return float(np.sum(ve) * scale - kl), kl


# --------------------------------------------------------------------------------------
# W initialiser (linear_svgp.py:17-48)
# --------------------------------------------------------------------------------------
def initialize_W(output_dim, num_latents, window_fraction=0.3, scale=0.5):
    W = np.zeros((output_dim, num_latents))
    window = max(int(output_dim * window_fraction), 2)
    stride = max(output_dim // (num_latents - 1), 1)
    for j in range(num_latents):
        c = min(int(j * stride), output_dim - 1)
        for i in range(output_dim):
            dist = abs(i - c)
            if dist < window / 2:
                W[i, j] = np.exp(-0.1 * dist)
    return W * scale


# --------------------------------------------------------------------------------------
# TF 2.10 Keras Adam + CosineDecay as driven by the reference loops
# --------------------------------------------------------------------------------------
class TFAdam:
    """ResourceApplyAdam with float32-stored hypers cast to float64."""

    def __init__(self, lr=0.001, beta1=0.9, beta2=0.999, eps=1e-7, cosine_decay_steps=None):
        self.lr0 = np.float32(lr)
        self.b1 = float(np.float32(beta1))
        self.b2 = float(np.float32(beta2))
        self.eps = float(eps)
        self.decay_steps = cosine_decay_steps
        self.t = 0
        self.m = None
        self.v = None

    def lr_at(self, step):
        if self.decay_steps is None:
            return float(self.lr0)
        s = np.float32(min(step, self.decay_steps))
        frac = s / np.float32(self.decay_steps)
        cosd = np.float32(0.5) * (np.float32(1.0) + np.cos(np.float32(math.pi) * frac, dtype=np.float32))
        return float(np.float32(self.lr0 * cosd))

    def step(self, params, grads):
        """params, grads: lists of float64 ndarrays (unconstrained); updates in place."""
        if self.m is None:
            self.m = [np.zeros_like(p) for p in params]
            self.v = [np.zeros_like(p) for p in params]
        lr = self.lr_at(self.t)
        self.t += 1
        t = float(self.t)
        lr_t = lr * math.sqrt(1.0 - self.b2**t) / (1.0 - self.b1**t)
        for p, g, m, v in zip(params, grads, self.m, self.v):
            m += (g - m) * (1.0 - self.b1)
            v += (g * g - v) * (1.0 - self.b2)
            p -= lr_t * m / (np.sqrt(v) + self.eps)


# --------------------------------------------------------------------------------------
# fixtures (committed under tests/golden/, generated by tests/golden/make_golden.py)
# --------------------------------------------------------------------------------------
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_dataset(name):
    """name in {"hbs", "goku"} -> dict with X [N, d+1], Y [N, P], X_test_aug, Y_test.

    Normalisation follows the reference drivers (tests/test_ho2021_multibin.py:24-35,
    data_loader.py:325-360, latin_hypercube.py:141-160).
    """
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    lim = z["input_limits"]

    def unit(x):
        return (x - lim[:, 0]) / (lim[:, 1] - lim[:, 0])

    XL, XH = unit(z["X_LF"]), unit(z["X_HF"])
    YL = z["Y_LF"] - z["Y_LF"].mean(axis=0)
    YH = z["Y_HF"]
    X = np.vstack([np.hstack([XL, np.zeros((XL.shape[0], 1))]), np.hstack([XH, np.ones((XH.shape[0], 1))])])
    Y = np.vstack([YL, YH])
    Xt = unit(z["X_test"])
    out = dict(X=X, Y=Y, X_test=np.hstack([Xt, np.ones((Xt.shape[0], 1))]), Y_test=z["Y_test"], kf=z["kf"])
    for k in z.files:
        if k.startswith("Z_kmeans"):
            out[k] = z[k]
    return out


def forrester_dataset():
    """tests/test_forrest.py:12-36 (draw order exactly as in the file)."""

    def forrester(x, sd=0):
        x = x.reshape((len(x), 1))
        f = ((6 * x - 2) ** 2) * np.sin(12 * x - 4)
        noise = np.random.normal(0, sd, x.shape) if sd > 0 else np.zeros_like(x)
        return f + noise

    def forrester_low(x, sd=0):
        return 0.5 * forrester(x, 0) + 10 * (x - 0.5) + 5 + np.random.randn(*x.shape) * sd

    state = np.random.get_state()
    np.random.seed(42)
    xl = np.random.rand(60, 1)
    xh = np.random.permutation(xl)[:20]
    yl = forrester_low(xl, sd=0.05)
    yh = forrester(xh, sd=0.02)
    np.random.set_state(state)
    X = np.vstack([np.hstack([xl, np.zeros_like(xl)]), np.hstack([xh, np.ones_like(xh)])])
    Y = np.vstack([yl, yh])
    xp = np.linspace(0, 1, 200)[:, None]
    return dict(X=X, Y=Y, X_plot_L=np.hstack([xp, np.zeros_like(xp)]), X_plot_H=np.hstack([xp, np.ones_like(xp)]))


def synthetic_exact_dataset(N, d=10, seed=0):
    """two-fidelity synthetic exact GPR."""
    rng = np.random.default_rng(seed)
    nH = N // 8
    nL = N - nH
    xL = rng.random((nL, d))
    xH = xL[rng.permutation(nL)[:nH]]
    fL = lambda x: np.sum(np.sin(2 * np.pi * x), axis=1)
    yL = fL(xL) + 0.03 * rng.standard_normal(nL)
    yH = 1.5 * fL(xH) + 0.3 * np.cos(2 * np.pi * xH[:, 0]) + 0.03 * rng.standard_normal(nH)
    X = np.vstack([np.hstack([xL, np.zeros((nL, 1))]), np.hstack([xH, np.ones((nH, 1))])])
    Y = np.concatenate([yL, yH])[:, None]
    theta = pack_theta(1.0, np.linspace(0.5, 1.0, d), 1.0, np.linspace(0.5, 1.0, d), 1.0)
    return dict(X=X, Y=Y, theta=theta, noise=1e-3)
