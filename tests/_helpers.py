"""Shared test helpers: oracle-side training steps in unconstrained space."""
import json
import os

import numpy as np

from oracle import mfgp_oracle as onp
from oracle import mfgp_oracle_torch as otc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def goldens():
    return json.load(open(os.path.join(GOLDEN, "goldens.json")))


def dsoftplus(theta, lower=0.0):
    """d theta / d u for theta = lower + softplus(u)."""
    return 1.0 - np.exp(-(np.asarray(theta) - lower))


def gpr_adam_trajectory(X, Y, theta0, noise, lr, n_steps, record_at):
    """Reference loop linear.py:203-221 with noise fixed (quirk Q3).  Returns {iter: LML}."""
    u = onp.softplus_inv(theta0).copy()
    opt = onp.TFAdam(lr=lr)
    out = {}
    for i in range(n_steps + 1):
        theta = onp.softplus(u)
        lml, g, _ = otc.gpr_lml_value_and_grad(X, Y, theta, noise)
        if i in record_at:
            out[i] = lml
        gu = -g * dsoftplus(theta)
        opt.step([u], [gu])
    return out


def svgp_one_adam_step(X, Y, Z, thetas, q_mu, q_sqrt, lik_var, W=None, num_data=None, lr=0.1, decay_steps=2000):
    """One optimisation step of singlebin_svgp.py:79-85 / linear_svgp.py:181-190, then -ELBO."""
    r = otc.svgp_value_and_grad(X, Y, Z, thetas, q_mu, q_sqrt, lik_var, W, num_data)
    u_th = onp.softplus_inv(thetas)
    u_lv = onp.softplus_inv(np.asarray(lik_var - onp.LIK_VAR_LOWER))
    params = [q_mu.copy(), q_sqrt.copy(), Z.copy(), u_th, np.atleast_1d(u_lv).astype(float)]
    grads = [r["g_q_mu"], r["g_q_sqrt"], r["g_Z"], r["g_thetas"] * dsoftplus(thetas),
             np.atleast_1d(r["g_lik_var"] * dsoftplus(lik_var, onp.LIK_VAR_LOWER))]
    if W is not None:
        params.append(W.copy())
        grads.append(r["g_W"])
    opt = onp.TFAdam(lr=lr, cosine_decay_steps=decay_steps)
    opt.step(params, grads)
    q_mu1, q_sqrt1, Z1, u_th1, u_lv1 = params[:5]
    W1 = params[5] if W is not None else None
    elbo1, _ = onp.svgp_elbo(X, Y, Z1, onp.softplus(u_th1), q_mu1, q_sqrt1,
                             float(onp.LIK_VAR_LOWER + onp.softplus(u_lv1[0])), W1, num_data)
    return r, -elbo1


class OracleSvgpOps:
    """CPU stand-in for dist.SvgpDeviceOps (HOST-LOGIC tests of the data-parallel loop only): the three per-step calls of
    include/mfgp.h's data-parallel SVGP section, computed by the torch oracle on CPU tensors."""

    def __init__(self):
        import torch

        self.device = torch.device("cpu")

    def make_cfg(self, L, M, P, B, d, hetero, scale, kl_mult, lik_lower, masked, lik_per_output, jitter=1e-6):
        from types import SimpleNamespace

        return SimpleNamespace(L=L, M=M, P=P, B=B, d=d, hetero=hetero, scale=scale, kl_mult=kl_mult, lik_lower=lik_lower,
                               masked=masked, lik_per_output=lik_per_output)

    @staticmethod
    def layout(cfg, has_W):
        n_theta = cfg.L * (2 * cfg.d + 3)
        o_Z = n_theta
        o_W = o_Z + cfg.M * (cfg.d + 1)
        o_qm = o_W + (cfg.P * cfg.L if has_W else 0)
        o_qs = o_qm + cfg.M * cfg.L
        o_lv = o_qs + cfg.L * cfg.M * cfg.M
        return n_theta, o_Z, o_W, o_qm, o_qs, o_lv, o_lv + (cfg.P if cfg.lik_per_output else 1)

    def constrain(self, cfg, has_W, u, c):
        n_theta, _, _, _, _, o_lv, n = self.layout(cfg, has_W)
        x = u.numpy()
        out = x.copy()
        out[:n_theta] = onp.softplus(x[:n_theta])
        out[o_lv:] = cfg.lik_lower + onp.softplus(x[o_lv:])
        c.copy_(__import__("torch").from_numpy(out))

    def elbo_grad_flat(self, cfg, X, Y, has_W, c, nranks, eg):
        n_theta, o_Z, o_W, o_qm, o_qs, o_lv, n = self.layout(cfg, has_W)
        v = c.numpy()
        L, M, P, d = cfg.L, cfg.M, cfg.P, cfg.d
        th, Z = v[:n_theta].reshape(L, 2 * d + 3), v[o_Z:o_W].reshape(M, d + 1)
        W = v[o_W:o_qm].reshape(P, L) if has_W else None
        q_mu, q_sqrt = v[o_qm:o_qs].reshape(M, L), v[o_qs:o_lv].reshape(L, M, M)
        lik = v[o_lv:] if cfg.lik_per_output else float(v[o_lv])
        r = otc.svgp_value_and_grad(X.numpy(), Y.numpy(), Z, th, q_mu, q_sqrt, lik, W, num_data=cfg.scale * X.shape[0],
                                    hetero=cfg.hetero, kl_mult=cfg.kl_mult / nranks, masked=cfg.masked)
        out = np.zeros(n + 2)
        out[0], out[1] = r["elbo"] + r["kl"], r["kl"] / nranks
        g = out[2:]
        g[:n_theta], g[o_Z:o_W] = r["g_thetas"].ravel(), r["g_Z"].ravel()
        if has_W:
            g[o_W:o_qm] = r["g_W"].ravel()
        g[o_qm:o_qs], g[o_qs:o_lv], g[o_lv:] = r["g_q_mu"].ravel(), r["g_q_sqrt"].ravel(), np.ravel(r["g_lik_var"])
        eg.copy_(__import__("torch").from_numpy(out))

    def adam_update(self, cfg, has_W, u, m, v, mask, c, eg, lr_t, step, b1, b2, eps, loss_hist, kl_hist, scratch):
        n_theta, _, _, _, _, o_lv, n = self.layout(cfg, has_W)
        s = int(step[0])
        e = eg.numpy()
        loss_hist[s] = -(e[0] - e[1]) + (cfg.kl_mult - 1.0) * e[1]
        kl_hist[s] = e[1]
        cc, gu = c.numpy(), e[2:].copy()
        gu[:n_theta] *= 1.0 - np.exp(-cc[:n_theta])
        gu[o_lv:] *= 1.0 - np.exp(-(cc[o_lv:] - cfg.lik_lower))
        on = np.ones(n, dtype=bool) if mask is None else mask.numpy().astype(bool)
        mm, vv, uu = m.numpy(), v.numpy(), u.numpy()
        mm[on] += (gu[on] - mm[on]) * (1.0 - b1)
        vv[on] += (gu[on] * gu[on] - vv[on]) * (1.0 - b2)
        uu[on] -= float(lr_t[s]) * mm[on] / (np.sqrt(vv[on]) + eps)
        step += 1
