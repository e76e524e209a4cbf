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

    def __init__(self, variance=1.0):
        self.variance = Parameter(variance, transform=positive(lower=DEFAULT_VARIANCE_LOWER_BOUND))


class HeteroscedasticGaussian(Gaussian):
    """Targets are [Y_obs | Y_unc]; effective variance = variance + Y_unc**2 (linear_svgp.py:259 --
    the code squares Y_unc although its docstring says otherwise; code wins, quirk Q6)."""

    heteroscedastic = True

    def __init__(self, variance):
        self.variance = Parameter(np.asarray(variance, dtype=np.float64), transform=positive())  # linear_svgp.py:240
