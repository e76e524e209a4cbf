"""Likelihood parameter holders: gpflow.likelihoods.Gaussian and the reference's
HeteroscedasticGaussian (mfgpflow/linear_svgp.py:223-267).  The variational expectations
themselves are a fused epilogue of the CUDA SVGP kernel sequence (csrc/svgp.cu)."""
from __future__ import annotations

import numpy as np

from .base import Parameter, positive

DEFAULT_VARIANCE_LOWER_BOUND = 1e-6  # gpflow.likelihoods.Gaussian


class Gaussian:
    _param_order = ("variance",)
    heteroscedastic = False
    masked = False

    def __init__(self, variance=1.0):
        self.variance = Parameter(variance, transform=positive(lower=DEFAULT_VARIANCE_LOWER_BOUND))


class HeteroscedasticGaussian(Gaussian):
    """Targets are [Y_obs | Y_unc]; effective variance = variance + Y_unc**2 (linear_svgp.py:259 --
    the code squares Y_unc although its docstring says otherwise; code wins, quirk Q6)."""

    heteroscedastic = True

    def __init__(self, variance):
        self.variance = Parameter(np.asarray(variance, dtype=np.float64), transform=positive())  # linear_svgp.py:240


class MaskedGaussian(Gaussian):
    """Gaussian likelihood that ignores the NaN entries of Y (missing outputs) in the variational expectations, with one
    trainable variance per output: reference notebooks/"demo: missing output.ipynb" cell 2 (class MaskedGaussian, built
    with variance=np.ones(P)).  Like every likelihood here it only holds parameters; the masking is the `masked` epilogue
    of the CUDA kernel (csrc/svgp.cu: mix_ve_kernel)."""

    masked = True

    def __init__(self, variance):
        super().__init__(np.atleast_1d(np.asarray(variance, dtype=np.float64)))
