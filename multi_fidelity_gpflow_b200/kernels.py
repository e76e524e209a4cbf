"""Kernel objects with the reference's surface: gpflow.kernels.SquaredExponential parameters and
LinearMultiFidelityKernel.K / K_diag (mfgpflow/linear.py:12-136).  All arithmetic is the CUDA
covariance kernel behind mfgp_cov / mfgp_cov_diag."""
from __future__ import annotations

from copy import deepcopy

import numpy as np

from . import _lib
from .base import Parameter, positive, set_trainable


class SquaredExponential:
    """gpflow.kernels.SquaredExponential / RBF parameter holder (ARD length-scales)."""

    _param_order = ("variance", "lengthscales")

    def __init__(self, variance=1.0, lengthscales=1.0):
        self.variance = Parameter(variance, transform=positive())
        self.lengthscales = Parameter(lengthscales, transform=positive())

    def ard(self, d):
        ls = np.atleast_1d(self.lengthscales.numpy())
        return np.full(d, ls[0]) if ls.size == 1 else ls

    def K(self, X, X2=None, handle=None):
        """Plain SE covariance through the same CUDA kernel (all points at fidelity 0)."""
        h = handle or _lib.default_handle()
        X = np.asarray(X, dtype=np.float64)
        d = X.shape[1]
        aug = lambda a: np.hstack([a, np.zeros((a.shape[0], 1))])
        theta = np.concatenate([[1.0], self.ard(d), [float(self.variance.numpy())], np.ones(d), [1.0]])
        return h.cov(aug(X), None if X2 is None else aug(np.asarray(X2, dtype=np.float64)), theta)

    def K_diag(self, X):
        return np.full(np.asarray(X).shape[0], float(self.variance.numpy()))


RBF = SquaredExponential


class LinearMultiFidelityKernel:
    """f_H(x) = rho f_L(x) + delta(x)   (Kennedy & O'Hagan 2000), reference linear.py:12-136.

    X carries the fidelity indicator in its last column.  `rho` has shape (num_output_dims, 1);
    like the reference (linear.py:90, quirk Q2) only rho[ith_output_dim=0] enters K.
    """

    _param_order = ("rho", "kernel_L", "kernel_delta")

    def __init__(self, kernel_L, kernel_delta, num_output_dims, use_rho=True, handle=None):
        self.kernel_L = kernel_L
        self.kernel_delta = kernel_delta
        self.rho = Parameter(np.ones((num_output_dims, 1)), transform=positive())
        if not use_rho:
            set_trainable(self.rho, False)  # linear.py:51-52
        self._handle = handle

    @property
    def handle(self):
        return self._handle or _lib.default_handle()

    def theta(self, d, ith_output_dim=0):
        """[rho, ls_L[d], var_L, ls_delta[d], var_delta] -- the layout of include/mfgp.h."""
        return np.concatenate([
            [float(self.rho.numpy()[ith_output_dim, 0])], self.kernel_L.ard(d), [float(self.kernel_L.variance.numpy())],
            self.kernel_delta.ard(d), [float(self.kernel_delta.variance.numpy())],
        ])

    def K(self, X, X2=None, ith_output_dim=0):
        X = np.asarray(X, dtype=np.float64)
        return self.handle.cov(X, X2, self.theta(X.shape[1] - 1, ith_output_dim))

    def K_diag(self, X, ith_output_dim=0):
        X = np.asarray(X, dtype=np.float64)
        return self.handle.cov_diag(X, self.theta(X.shape[1] - 1, ith_output_dim))

    def __call__(self, X, X2=None, full_cov=True):
        return self.K(X, X2) if full_cov else self.K_diag(X)

    # ---- gradient plumbing: constrained d/dtheta (layout above) -> per-Parameter unconstrained ----
    def scatter_theta_grad(self, g_theta, d, ith_output_dim=0):
        """Returns {Parameter: unconstrained gradient} for one theta-gradient vector [2d+3]."""
        g_rho = np.zeros(self.rho.shape)
        g_rho[ith_output_dim, 0] = g_theta[0]
        out = [(self.rho, g_rho)]
        for kern, sl, iv in ((self.kernel_L, slice(1, 1 + d), 1 + d), (self.kernel_delta, slice(2 + d, 2 + 2 * d), 2 + 2 * d)):
            gl = np.asarray(g_theta[sl])
            if kern.lengthscales.shape == ():  # isotropic length-scale shared by all dimensions
                gl = np.sum(gl)
            out.append((kern.lengthscales, gl))
            out.append((kern.variance, g_theta[iv]))
        return [(p, p.grad_to_unconstrained(g)) for p, g in out]

    def trainable_parameters(self):
        """GPflow order seen in the notebooks: rho, kernel_L.lengthscales, kernel_L.variance, kernel_delta.*"""
        ps = [self.rho, self.kernel_L.lengthscales, self.kernel_L.variance, self.kernel_delta.lengthscales,
              self.kernel_delta.variance]
        return [p for p in ps if p.trainable]


class SeparateIndependent:
    """gpflow.kernels.SeparateIndependent: one independent kernel per output (singlebin_svgp.py:47)."""

    _param_order = ("kernels",)

    def __init__(self, kernels):
        self.kernels = list(kernels)
        self.W = None

    @property
    def num_latent_gps(self):
        return len(self.kernels)


class LinearCoregionalization:
    """gpflow.kernels.LinearCoregionalization: f = W g (linear_svgp.py:122)."""

    _param_order = ("kernels", "W")

    def __init__(self, kernels, W):
        self.kernels = list(kernels)
        self.W = W if isinstance(W, Parameter) else Parameter(W)

    @property
    def num_latent_gps(self):
        return len(self.kernels)


def replicate_mf_kernels(kernel_L, kernel_delta, count, use_rho=True, handle=None):
    """[LinearMultiFidelityKernel(deepcopy(kernel_L), deepcopy(kernel_delta), 1) ...] (singlebin_svgp.py:39)."""
    return [LinearMultiFidelityKernel(deepcopy(kernel_L), deepcopy(kernel_delta), 1, use_rho=use_rho, handle=handle)
            for _ in range(count)]
