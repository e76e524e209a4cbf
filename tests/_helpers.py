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
