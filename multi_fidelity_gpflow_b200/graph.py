"""GraphMultiFidelityKernel / GraphMultiFidelityGPModel -- host-side mirror of the reference's graph-structured
multi-fidelity model (mfgpflow/graph.py:7-188): several low-fidelity sources with learnable cross-correlations,
f_H(x) = sum_i rho_i f_Li(x) + delta(x).  Same constructor arguments, attributes (`rho` (num_LF, P), `rho_LF`
(num_LF, num_LF) with a sigmoid transform, `kernel_Ls`, `kernel_delta`), `K`, `K_diag`, `log_marginal_likelihood`,
`training_loss`, `trainable_variables` and `optimize` (Adam | SciPy L-BFGS-B) as the reference; covariance, objective
and analytic gradient run in libmfgp.so (csrc/graph.cu).

Scope: the symmetric call K(X) and the training objective.  The reference's rectangular
K(X, X2) is not shape-consistent (graph.py:76-79 scatters a |H2| x |L_i| block into an |H| x |L2_i| index grid and :91 adds
eye(N) to an N x N2 matrix), so `predict_f` -- which needs K(X, Xnew) -- has no defined reference behaviour and raises."""
from __future__ import annotations

import numpy as np

from . import _lib
from .base import Parameter, positive, set_trainable, sigmoid
from .likelihoods import Gaussian
from .optimizers import Adam, Scipy


class GraphMultiFidelityKernel:
    _param_order = ("rho", "rho_LF", "kernel_Ls", "kernel_delta")

    def __init__(self, kernel_Ls, kernel_delta, num_LF, num_output_dims, handle=None):
        if len(kernel_Ls) != num_LF:
            raise ValueError("one low-fidelity kernel per source")
        self.num_LF = int(num_LF)
        self.kernel_Ls = list(kernel_Ls)
        self.kernel_delta = kernel_delta
        self.rho = Parameter(np.ones((num_LF, num_output_dims)), transform=positive())  # graph.py:30-32
        self.rho_LF = Parameter(0.5 * np.ones((num_LF, num_LF)), transform=sigmoid())    # graph.py:35-37
        self._handle = handle

    @property
    def handle(self):
        return self._handle or _lib.default_handle()

    def gtheta(self, d, ith_output_dim=0):
        """[rho (m), rho_LF (m x m), (ls_Li (d), var_Li) for every source, ls_delta (d), var_delta]: include/mfgp.h layout."""
        parts = [self.rho.numpy()[:, ith_output_dim], self.rho_LF.numpy().ravel()]
        for k in self.kernel_Ls + [self.kernel_delta]:
            parts += [k.ard(d), [float(k.variance.numpy())]]
        return np.concatenate([np.ravel(p) for p in parts]).astype(np.float64)

    def K(self, X, X2=None, ith_output_dim=0):
        if X2 is not None and X2 is not X:
            raise NotImplementedError("GraphMultiFidelityKernel.K(X, X2): the reference is only shape-consistent for X2 = X")
        X = np.asarray(X, dtype=np.float64)
        return self.handle.graph_cov(X, self.num_LF, self.gtheta(X.shape[1] - 1, ith_output_dim))

    def K_diag(self, X, ith_output_dim=0):
        X = np.asarray(X, dtype=np.float64)
        return self.handle.graph_cov_diag(X, self.num_LF, self.gtheta(X.shape[1] - 1, ith_output_dim))

    def scatter_gtheta_grad(self, g, d, ith_output_dim=0):
        """[(Parameter, unconstrained gradient)] from one constrained gradient vector in the gtheta layout."""
        m = self.num_LF
        g_rho = np.zeros(self.rho.shape)
        g_rho[:, ith_output_dim] = g[:m]  # only column ith_output_dim enters K (graph.py:51; GPflow always passes 0)
        g_rlf = np.asarray(g[m:m + m * m]).reshape(m, m).copy()
        g_rlf[np.diag_indices(m)] = 0.0   # rho_LF[i, i] is never read (graph.py:62)
        out = [(self.rho, g_rho), (self.rho_LF, g_rlf)]
        o = m + m * m
        for k in self.kernel_Ls + [self.kernel_delta]:
            gl = np.asarray(g[o:o + d])
            out.append((k.lengthscales, np.sum(gl) if k.lengthscales.shape == () else gl))
            out.append((k.variance, g[o + d]))
            o += d + 1
        return [(p, p.grad_to_unconstrained(v)) for p, v in out]

    def trainable_parameters(self):
        ps = [self.rho, self.rho_LF]
        for k in self.kernel_Ls + [self.kernel_delta]:
            ps += [k.lengthscales, k.variance]
        return [p for p in ps if p.trainable]


class GraphMultiFidelityGPModel:
    _param_order = ("kernel", "likelihood")

    def __init__(self, X, Y, kernel_Ls, kernel_delta, handle=None):
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        self._handle = handle
        self.num_LF = len(kernel_Ls)
        self.num_output_dims = Y.shape[1]
        self.kernel = GraphMultiFidelityKernel(kernel_Ls, kernel_delta, self.num_LF, self.num_output_dims, handle=handle)
        self.likelihood = Gaussian(variance=1e-3)       # graph.py:135
        set_trainable(self.likelihood.variance, False)  # graph.py:138
        self.data = (X, Y)
        self.loss_history = []

    @property
    def handle(self):
        return self._handle or _lib.default_handle()

    def _call(self, want_grad):
        X, Y = self.data
        return self.handle.graph_gpr_nlml_grad(X, Y, self.num_LF, self.kernel.gtheta(X.shape[1] - 1),
                                               float(self.likelihood.variance.numpy()), want_grad=want_grad)

    def log_marginal_likelihood(self):
        return -self._call(False)[0]

    def training_loss(self):
        return -self.log_marginal_likelihood()

    @property
    def trainable_variables(self):
        vs = self.kernel.trainable_parameters()
        if self.likelihood.variance.trainable:
            vs.append(self.likelihood.variance)
        return vs

    def value_and_grad(self, variables=None):
        variables = self.trainable_variables if variables is None else variables
        X, _ = self.data
        d = X.shape[1] - 1
        nlml, g = self._call(True)
        by = {id(p): gu for p, gu in self.kernel.scatter_gtheta_grad(g[:-1], d)}
        by[id(self.likelihood.variance)] = self.likelihood.variance.grad_to_unconstrained(g[-1])
        return nlml, [by[id(p)] for p in variables]

    def predict_f(self, Xnew, full_cov=False, full_output_cov=False):
        raise NotImplementedError("GraphMultiFidelityKernel.K(X, Xnew) is not shape-consistent in the reference "
                                  "(graph.py:76-79, :91); only the training objective is defined")

    def optimize(self, max_iters=1000, learning_rate=0.01, use_adam=True, unfix_noise_after=500, verbose=True):
        """graph.py:144-188: same two loops as MultiFidelityGPModel.optimize, including the no-op noise un-fix under Adam
        (the tf.function captured `trainable_variables` once, quirk Q3)."""
        self.loss_history = []
        if use_adam:
            optimizer = Adam(learning_rate)
            traced = self.trainable_variables
            for i in range(max_iters):
                loss, grads = self.value_and_grad(traced)
                optimizer.apply_gradients(zip(grads, traced))
                self.loss_history.append(loss)
                if i == unfix_noise_after:
                    set_trainable(self.likelihood.variance, True)
                if verbose and i % 100 == 0:
                    print(f"Iteration {i}: Loss = {-loss}")
        else:
            opt = Scipy()
            for _ in range(2):
                vs = self.trainable_variables

                def closure(vs=vs):
                    loss, grads = self.value_and_grad(vs)
                    self.loss_history.append(loss)
                    return loss, grads

                opt.minimize(closure, vs, options={"maxiter": max_iters})
                set_trainable(self.likelihood.variance, True)
        return self
