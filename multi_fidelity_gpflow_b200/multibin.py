"""One linear multi-fidelity GPR per output bin, many hyper-parameter restarts per bin, trained together on the device.

This is BASELINE config 2 ("Ho-Bird-Shelton 2021 50LF-3HF multi-bin: one linear MF GPR per k-bin, batched"), the
"many single-output GP" design announced in reference `mfgpflow/gpemulator_singlebin.py:1-14`, driven by the Adam branch of
`MultiFidelityGPModel.optimize` (`mfgpflow/linear.py:190-221`: loss = -log_marginal_likelihood, gradient w.r.t. the
unconstrained variables, Keras Adam with float32-stored hyper-parameters, likelihood noise fixed at 1e-3 -- quirk Q3).
Every bin p and restart r is an independent problem b = r * P + p of `mfgp_gpr_batched_adam`: the whole
loop -- softplus, K6 NLML + gradient, chain rule, Adam -- runs on the GPU; the host only supplies the per-step factors
lr(step) * sqrt(1 - beta2^t) / (1 - beta1^t).  No CPU fallback: every number comes from libmfgp.so.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .base import positive
from .optimizers import adam_step_factors


class MultiBinMFGP:
    def __init__(self, X, Y, num_restarts=1, noise=1e-3, seed=0, spread=0.3, use_rho=True, handle=None):
        """X [N, d+1] with the fidelity column last, Y [N, P] (one column per bin), N <= 64.  Restart 0 starts at the
        reference's initial values (rho = 1, lengthscales = 1, variances = 1: linear.py:42-52 with gpflow defaults);
        restarts r > 0 at log-normal perturbations of them (sigma = `spread`, seeded)."""
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.Y = np.ascontiguousarray(Y, dtype=np.float64)
        self.N, self.d = self.X.shape[0], self.X.shape[1] - 1
        self.P, self.R = self.Y.shape[1], int(num_restarts)
        if self.N > 64:
            raise ValueError("MultiBinMFGP runs the small-matrix kernel (N <= 64); use MultiFidelityGPModel for larger problems")
        self.handle = handle if handle is not None else _lib.default_handle()
        self.noise = float(noise)
        self.use_rho = bool(use_rho)
        self._tf = positive()
        np_ = 2 * self.d + 3
        rng = np.random.default_rng(seed)
        theta0 = np.ones((self.R, self.P, np_))
        if self.R > 1:
            theta0[1:] *= np.exp(spread * rng.standard_normal((self.R - 1, self.P, np_)))
        self.u = np.ascontiguousarray(self._tf.inverse(theta0).reshape(self.R * self.P, np_))
        self.m = np.zeros_like(self.u)
        self.v = np.zeros_like(self.u)
        self.iterations = 0
        self.loss_history = np.zeros((0, self.R, self.P))
        self.failed = np.zeros((self.R, self.P), dtype=np.int32)  # first failing Cholesky pivot per (restart, bin), 0 = healthy

    # ---- parameters --------------------------------------------------------------------------------
    @property
    def thetas(self):
        """Constrained hyper-parameters [R, P, 2d+3] = [rho, ls_L (d), var_L, ls_delta (d), var_delta]."""
        return self._tf.forward(self.u).reshape(self.R, self.P, -1)

    def training_loss(self):
        """NLML of every (restart, bin): [R, P]; NaN where that problem's covariance is not positive definite."""
        B = self.R * self.P
        info = np.zeros(B, dtype=np.int32)
        nlml, _ = self.handle.gpr_batched_nlml_grad(self.X, self.Y, self.thetas.reshape(B, -1), np.full(B, self.noise),
                                                    info=info, want_grad=False)
        return np.where(info == 0, nlml, np.nan).reshape(self.R, self.P)

    def log_marginal_likelihood(self):
        return -self.training_loss()

    # ---- training ------------------------------------------------------------------------------------
    def optimize(self, max_iters=1000, learning_rate=0.01, use_cosine_decay=False, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        """max_iters Adam steps for all R * P problems on the device.  A restart whose covariance stops being positive
        definite is that restart's failure only: `failed[r, p]` records the 1-based pivot of its first failure, its later
        losses are NaN (so best_restart() skips it) and the remaining restarts train on -- the step counter and
        loss_history advance for everyone, so the bias correction / schedule stay right on the next call."""
        lr_t, b1, b2 = adam_step_factors(learning_rate, max_iters, first_step=self.iterations,
                                         cosine_decay_steps=max_iters if use_cosine_decay else None, beta_1=beta_1, beta_2=beta_2)
        B = self.R * self.P
        hist = np.empty((max_iters, B))
        info = np.zeros(B, dtype=np.int32)
        self.handle.gpr_batched_adam(self.X, self.Y, self.u, self.m, self.v, np.full(B, self.noise), lr_t, b1, b2, epsilon,
                                     fix_rho=not self.use_rho, loss_hist=hist, info=info)
        self.iterations += max_iters
        self.loss_history = np.concatenate([self.loss_history, hist.reshape(max_iters, self.R, self.P)])
        new = info.reshape(self.R, self.P)
        self.failed = np.where(self.failed != 0, self.failed, new)
        return self

    # ---- model selection and prediction ----------------------------------------------------------------
    def best_restart(self):
        """Per bin, the restart with the lowest NLML at the current hyper-parameters: [P]."""
        loss = self.training_loss()
        return np.argmin(np.where(np.isfinite(loss), loss, np.inf), axis=0)

    def best_thetas(self):
        r = self.best_restart()
        return self.thetas[r, np.arange(self.P)]

    def predict_f(self, Xnew):
        """Posterior mean / variance [N*, P] of every bin at its best restart (GPR.predict_f per bin)."""
        Xnew = np.ascontiguousarray(Xnew, dtype=np.float64)
        th = self.best_thetas()
        mean, var = np.empty((Xnew.shape[0], self.P)), np.empty((Xnew.shape[0], self.P))
        for p in range(self.P):
            mu, s2 = self.handle.gpr_predict(self.X, np.ascontiguousarray(self.Y[:, p:p + 1]), Xnew, th[p], self.noise)
            mean[:, p], var[:, p] = mu[:, 0], s2
        return mean, var
