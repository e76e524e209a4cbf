"""SingleBinSVGP -- mirror of mfgpflow/singlebin_svgp.py:13-135: P independent multi-fidelity
kernels (SeparateIndependent), shared KMeans inducing points, whitened SVGP, Adam + cosine decay."""
from __future__ import annotations

import numpy as np

from .base import set_trainable
from .kernels import SeparateIndependent, replicate_mf_kernels
from .likelihoods import Gaussian
from .optimizers import Adam, CosineDecay
from .svgp_base import SVGPBase, kmeans_inducing_points


class SingleBinSVGP(SVGPBase):
    def __init__(self, X, Y, kernel_L, kernel_delta, num_outputs, Z, random_state=42, handle=None):
        self.num_outputs = num_outputs
        kernel = SeparateIndependent(replicate_mf_kernels(kernel_L, kernel_delta, num_outputs, handle=handle))  # :39,47
        M = np.asarray(Z).shape[0]
        Z_init = kmeans_inducing_points(np.asarray(X, dtype=np.float64), M, random_state)  # :50-51
        q_mu = np.zeros((M, num_outputs))  # :56
        q_sqrt = np.repeat(np.eye(M)[None, ...], num_outputs, axis=0) * 0.1  # :57
        # Gaussian() variance 1.0, trainable from step 0; no num_data => ELBO scale 1 (quirk Q4)
        self._init_svgp(kernel, Gaussian(), Z_init, num_outputs, q_mu, q_sqrt, num_data=None, handle=handle)

    def optimize(self, data, max_iters=10000, initial_lr=0.01, unfix_noise_after=5000, verbose=True, print_every=10):
        optimizer = Adam(CosineDecay(initial_lr, max_iters))
        self.loss_history = []
        traced = self.trainable_variables
        for i in range(max_iters):
            loss, _, grads = self.value_and_grad(data, traced)
            optimizer.apply_gradients(zip(grads, traced))
            self.loss_history.append(loss)
            if i == unfix_noise_after:
                set_trainable(self.likelihood.variance, True)  # redundant in the reference too (Q4)
            if verbose and i % print_every == 0:
                print(f"Iteration {i}: ELBO = {-self.elbo(data)}")  # the reference prints -ELBO after the step (:96-97)
        return self

    @staticmethod
    def load_model(filename, X, Y, kernel_L, kernel_delta, num_outputs, Z, *_ignored):
        # the reference's own test passes one extra positional argument (tests/test_ho2021_singlebin.py:146)
        return SingleBinSVGP(X, Y, kernel_L, kernel_delta, num_outputs, Z)._load_params(filename)
