"""torch-float64 CPU autograd twin of ``mfgp_oracle.py`` -- gradients for the parity tests.

TEST INFRASTRUCTURE ONLY (same rule as mfgp_oracle.py: never imported by the product).

The reference obtains gradients with ``tf.GradientTape`` through the graph restated in
``mfgp_oracle.py`` (``linear.py:205-207``, ``singlebin_svgp.py:82-84``,
``linear_svgp.py:183-189``).  This file builds the same graph in torch so reverse-mode
autodiff yields the same derivative; tests/test_oracle_fd.py checks it against central finite
differences of the NumPy forward (every parameter group, every likelihood variant).  It is also the "port" CPU baseline timed by bench.py
(all host threads), because the reference's TensorFlow/GPflow cannot be installed here.
"""
from __future__ import annotations

import math

import numpy as np
import torch

JITTER = 1e-6
LOG2PI = math.log(2.0 * math.pi)
DT = torch.float64


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x), dtype=DT)


def se_K(A, B, ls, var):
    A = A / ls
    B = B / ls
    As = (A * A).sum(-1)
    Bs = (B * B).sum(-1)
    dist = -2.0 * (A @ B.T) + As[:, None] + Bs[None, :]
    return var * torch.exp(-0.5 * dist)


def split_theta(theta, d):
    return theta[0], theta[1 : 1 + d], theta[1 + d], theta[2 + d : 2 + 2 * d], theta[2 + 2 * d]


def mf_K(X, X2, theta):
    """Same values as the gather/scatter form of linear.py:55-104, written with masks so
    that rows with fidelity not in {0, 1} come out exactly zero (quirk Q1)."""
    X = _t(X)
    X2 = X if X2 is None else _t(X2)
    d = X.shape[1] - 1
    rho, lsL, vL, lsD, vD = split_theta(theta, d)
    f, f2 = X[:, -1], X2[:, -1]
    one = torch.ones((), dtype=DT)
    zero = torch.zeros((), dtype=DT)
    s = torch.where(f == 0, one, torch.where(f == 1, rho, zero))
    s2 = torch.where(f2 == 0, one, torch.where(f2 == 1, rho, zero))
    h = (f == 1).to(DT)
    h2 = (f2 == 1).to(DT)
    x, x2 = X[:, :-1], X2[:, :-1]
    # NaN fidelity rows must not poison the product: zero the coordinates of dead rows
    live = ((f == 0) | (f == 1))[:, None]
    live2 = ((f2 == 0) | (f2 == 1))[:, None]
    x = torch.where(live, x, torch.zeros_like(x))
    x2 = torch.where(live2, x2, torch.zeros_like(x2))
    return (s[:, None] * s2[None, :]) * se_K(x, x2, lsL, vL) + (h[:, None] * h2[None, :]) * se_K(x, x2, lsD, vD)


def mf_K_diag(X, theta):
    X = _t(X)
    d = X.shape[1] - 1
    rho, _, vL, _, vD = split_theta(theta, d)
    f = X[:, -1]
    return (f == 0).to(DT) * vL + (f == 1).to(DT) * (rho * rho * vL + vD)


def gpr_lml(X, Y, theta, noise):
    X, Y = _t(X), _t(Y)
    N = X.shape[0]
    Kn = mf_K(X, None, theta) + noise * torch.eye(N, dtype=DT)
    L = torch.linalg.cholesky(Kn)
    A = torch.linalg.solve_triangular(L, Y, upper=False)
    P = Y.shape[1]
    return -0.5 * (A * A).sum() - P * (0.5 * N * LOG2PI + torch.log(torch.diagonal(L)).sum())


def gpr_lml_value_and_grad(X, Y, theta, noise):
    """Returns (lml, dlml/dtheta [2d+3], dlml/dnoise) w.r.t. CONSTRAINED parameters."""
    th = torch.tensor(np.asarray(theta, dtype=np.float64), requires_grad=True)
    nz = torch.tensor(float(noise), dtype=DT, requires_grad=True)
    val = gpr_lml(X, Y, th, nz)
    gth, gnz = torch.autograd.grad(val, [th, nz])
    return float(val.detach()), gth.numpy().copy(), float(gnz)


def gpr_batched_value_and_grad(X, Y, thetas, noises):
    """Independent GP per column (north-star multi-bin).  Returns lml [B], grad [B, 2d+4]."""
    B = Y.shape[1]
    vals = np.empty(B)
    grads = np.empty((B, thetas.shape[1] + 1))
    for b in range(B):
        v, g, gn = gpr_lml_value_and_grad(X, Y[:, b : b + 1], thetas[b], noises[b])
        vals[b] = v
        grads[b, :-1] = g
        grads[b, -1] = gn
    return vals, grads


def graph_K(X, gtheta, m):
    """GraphMultiFidelityKernel.K(X) (graph.py:39-93) written with masks; same values as oracle.mfgp_oracle.graph_K."""
    X = _t(X)
    N, d = X.shape[0], X.shape[1] - 1
    rho, rho_LF = gtheta[:m], gtheta[m:m + m * m].reshape(m, m)
    o = m + m * m
    f = X[:, -1]
    live = torch.zeros(N, dtype=torch.bool)
    for i in range(m + 1):
        live = live | (f == i)
    x = torch.where(live[:, None], X[:, :-1], torch.zeros_like(X[:, :-1]))
    ind = [(f == i).to(DT) for i in range(m + 1)]
    kern = []
    for i in range(m + 1):
        kern.append(se_K(x, x, gtheta[o:o + d], gtheta[o + d]))
        o += d + 1
    K = torch.zeros((N, N), dtype=DT)
    H = ind[m]
    for i in range(m):
        for j in range(m):
            coef = rho_LF[i, j] if i != j else torch.ones((), dtype=DT)
            K = K + coef * (ind[i][:, None] * ind[j][None, :]) * kern[i]
        K = K + rho[i] * (ind[i][:, None] * H[None, :] + H[:, None] * ind[i][None, :]) * kern[i]
        K = K + rho[i] ** 2 * (H[:, None] * H[None, :]) * kern[i]
    K = K + (H[:, None] * H[None, :]) * kern[m]
    return K + 1e-6 * torch.eye(N, dtype=DT)


def graph_gpr_lml_value_and_grad(X, Y, gtheta, m, noise):
    """LML of GraphMultiFidelityGPModel and its gradient w.r.t. the CONSTRAINED [gtheta, noise] exactly as
    tape.gradient produces it: the (possibly asymmetric) K goes into the Cholesky as is -- torch.linalg.cholesky, like
    tf.linalg.cholesky, reads the lower triangle and its backward returns the symmetrised sensitivity."""
    X, Y = _t(X), _t(Y)
    N, P = Y.shape
    th = torch.tensor(np.asarray(gtheta, dtype=np.float64), requires_grad=True)
    nz = torch.tensor(float(noise), dtype=DT, requires_grad=True)
    Kn = graph_K(X, th, m) + nz * torch.eye(N, dtype=DT)
    L = torch.linalg.cholesky(Kn)
    A = torch.linalg.solve_triangular(L, Y, upper=False)
    val = -0.5 * (A * A).sum() - P * (0.5 * N * LOG2PI + torch.log(torch.diagonal(L)).sum())
    gth, gnz = torch.autograd.grad(val, [th, nz])
    return float(val.detach()), gth.numpy().copy(), float(gnz)


def prior_kl(q_mu, q_sqrt):
    M, L = q_mu.shape
    Lq = torch.tril(q_sqrt)
    diag = torch.diagonal(Lq, dim1=-2, dim2=-1)
    return 0.5 * ((q_mu * q_mu).sum() - M * L - torch.log(diag * diag).sum() + (Lq * Lq).sum())


def svgp_predict(Xb, Z, thetas, q_mu, q_sqrt, W=None):
    Xb, Z = _t(Xb), _t(Z)
    Lnum, M = thetas.shape[0], Z.shape[0]
    gm, gv = [], []
    eye = torch.eye(M, dtype=DT)
    for l in range(Lnum):
        Kmm = mf_K(Z, None, thetas[l]) + JITTER * eye
        Kmn = mf_K(Z, Xb, thetas[l])
        knn = mf_K_diag(Xb, thetas[l])
        Lm = torch.linalg.cholesky(Kmm)
        A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
        LTA = torch.tril(q_sqrt[l]).T @ A
        gv.append(knn - (A * A).sum(0) + (LTA * LTA).sum(0))
        gm.append(A.T @ q_mu[:, l])
    g_mean = torch.stack(gm, dim=1)
    g_var = torch.stack(gv, dim=1)
    if W is None:
        return g_mean, g_var
    return g_mean @ W.T, g_var @ (W * W).T


def svgp_elbo(Xb, Yb, Z, thetas, q_mu, q_sqrt, lik_var, W=None, num_data=None, hetero=False, masked=False):
    Xb, Yb = _t(Xb), _t(Yb)
    kl = prior_kl(q_mu, q_sqrt)
    f_mean, f_var = svgp_predict(Xb, Z, thetas, q_mu, q_sqrt, W)
    P = f_mean.shape[1]
    if masked:  # MaskedGaussian (notebooks/"demo: missing output.ipynb" cell 2): tf.where fills, then zeroes the masked VE
        mask = ~torch.isnan(Yb)
        zero = torch.zeros_like(f_mean)
        Yf, Fm, Fv = torch.where(mask, Yb, zero), torch.where(mask, f_mean, zero), torch.where(mask, f_var, zero)
        ve = -0.5 * LOG2PI - 0.5 * torch.log(lik_var) - 0.5 * ((Yf - Fm) ** 2 + Fv) / lik_var
        ve = torch.where(mask, ve, zero)
        scale = 1.0 if num_data is None else float(num_data) / Xb.shape[0]
        return ve.sum() * scale - kl, kl
    if hetero:
        Yo, Yu = Yb[:, :P], Yb[:, P:]
        ev = lik_var + Yu * Yu
    else:
        Yo = Yb
        ev = lik_var * torch.ones_like(f_mean)
    ve = -0.5 * LOG2PI - 0.5 * torch.log(ev) - 0.5 * ((Yo - f_mean) ** 2 + f_var) / ev
    scale = 1.0 if num_data is None else float(num_data) / Xb.shape[0]
    return ve.sum() * scale - kl, kl


def svgp_value_and_grad(Xb, Yb, Z, thetas, q_mu, q_sqrt, lik_var, W=None, num_data=None, hetero=False, kl_mult=1.0,
                        masked=False):
    """loss = -ELBO + (kl_mult-1)*KL (linear_svgp.py:188).  Gradients w.r.t. CONSTRAINED values.

    Returns dict(loss, elbo, kl, g_Z, g_thetas, g_q_mu, g_q_sqrt, g_W, g_lik_var).
    """
    tz = torch.tensor(np.asarray(Z, dtype=np.float64), requires_grad=True)
    tth = torch.tensor(np.asarray(thetas, dtype=np.float64), requires_grad=True)
    tqm = torch.tensor(np.asarray(q_mu, dtype=np.float64), requires_grad=True)
    tqs = torch.tensor(np.asarray(q_sqrt, dtype=np.float64), requires_grad=True)
    lv = np.asarray(lik_var, dtype=np.float64)
    tlv = torch.tensor(lv.reshape(()) if lv.size == 1 else lv.ravel(), requires_grad=True)  # [P] vector: MaskedGaussian
    leaves = [tz, tth, tqm, tqs, tlv]
    tw = None
    if W is not None:
        tw = torch.tensor(np.asarray(W, dtype=np.float64), requires_grad=True)
        leaves.append(tw)
    elbo, kl = svgp_elbo(Xb, Yb, tz, tth, tqm, tqs, tlv, tw, num_data, hetero, masked)
    loss = -elbo + (kl_mult - 1.0) * kl
    gs = torch.autograd.grad(loss, leaves)
    out = dict(
        loss=float(loss.detach()),
        elbo=float(elbo.detach()),
        kl=float(kl.detach()),
        g_Z=gs[0].numpy().copy(),
        g_thetas=gs[1].numpy().copy(),
        g_q_mu=gs[2].numpy().copy(),
        g_q_sqrt=np.tril(gs[3].numpy()).copy(),
        g_lik_var=float(gs[4]) if lv.size == 1 else gs[4].numpy().copy(),
        g_W=None if W is None else gs[5].numpy().copy(),
    )
    return out
