"""Minimal loader for the reference's in-repo power-spectrum emulation arrays (the txt layout read
by mfgpflow/data_loader.py:288-322) and the normalisation every reference driver applies
(data_loader.py:325-360, latin_hypercube.py:141-160, tests/test_ho2021_multibin.py:24-35)."""
from __future__ import annotations

import os

import numpy as np


class PowerSpecs:
    """X_train / Y_train lists per fidelity, parameter_limits, X_test / Y_test, kf."""

    def __init__(self, folder: str | None = None, n_fidelities: int = 2):
        self.n_fidelities = n_fidelities
        if folder is not None:
            self.read_from_txt(folder)

    def read_from_txt(self, folder: str):
        f = lambda n: np.loadtxt(os.path.join(folder, n))
        self.X_train = [f(f"train_input_fidelity_{i}.txt") for i in range(self.n_fidelities)]
        self.Y_train = [f(f"train_output_fidelity_{i}.txt") for i in range(self.n_fidelities)]
        self.parameter_limits = f("input_limits.txt")
        self.X_test = [np.atleast_2d(f("test_input.txt"))]
        self.Y_test = [np.atleast_2d(f("test_output.txt"))]
        self.kf = f("kf.txt")
        assert len(self.kf) == self.Y_test[0].shape[1] == self.Y_train[0].shape[1]
        return self

    def read_from_npz(self, path: str):
        z = np.load(path)
        self.X_train = [z["X_LF"], z["X_HF"]]
        self.Y_train = [z["Y_LF"], z["Y_HF"]]
        self.parameter_limits = z["input_limits"]
        self.X_test, self.Y_test, self.kf = [z["X_test"]], [z["Y_test"]], z["kf"]
        self.extras = {k: z[k] for k in z.files if k.startswith("Z_kmeans")}
        return self

    def _unit(self, x):
        lim = self.parameter_limits
        return (x - lim[:, 0]) / (lim[:, 1] - lim[:, 0])

    @property
    def X_train_norm(self):
        return [self._unit(x) for x in self.X_train]

    @property
    def X_test_norm(self):
        return [self._unit(x) for x in self.X_test]

    @property
    def Y_train_norm(self):
        out = [y - y.mean(axis=0) for y in self.Y_train[:-1]]
        out.append(self.Y_train[-1])
        return out

    def training_arrays(self):
        """(X [N, d+1] with fidelity column, Y [N, P]) exactly as the reference drivers stack them."""
        Xs, Ys = self.X_train_norm, self.Y_train_norm
        X = np.vstack([np.hstack([x, np.full((x.shape[0], 1), float(i))]) for i, x in enumerate(Xs)])
        return X, np.vstack(Ys)

    def test_arrays(self, fidelity: float = 1.0):
        xt = self.X_test_norm[0]
        return np.hstack([xt, np.full((xt.shape[0], 1), fidelity)]), self.Y_test[0]


def synthetic_two_fidelity(N, d=10, seed=0):
    """BASELINE config 5 / SURVEY 8(d) C5: synthetic two-fidelity exact-GPR problem of N points in d dimensions.  The first
    7N/8 rows are low fidelity at uniform random inputs, the last N/8 high fidelity at a random subset of them (nested
    designs, like the in-repo datasets); f_L = sum_k sin(2 pi x_k), f_H = 1.5 f_L + 0.3 cos(2 pi x_0), noise sd 0.03.
    Returns (X [N, d+1], Y [N, 1], theta [2d+3], noise) with the configuration's hyper-parameters: lengthscales
    linspace(0.5, 1, d) for both kernels, unit variances, rho = 1, noise variance 1e-3."""
    rng = np.random.default_rng(seed)
    n_hi = N // 8
    n_lo = N - n_hi
    x_lo = rng.random((n_lo, d))
    x_hi = x_lo[rng.permutation(n_lo)[:n_hi]]
    wave = lambda x: np.sin(2.0 * np.pi * x).sum(axis=1)
    y_lo = wave(x_lo) + 0.03 * rng.standard_normal(n_lo)
    y_hi = 1.5 * wave(x_hi) + 0.3 * np.cos(2.0 * np.pi * x_hi[:, 0]) + 0.03 * rng.standard_normal(n_hi)
    X = np.zeros((N, d + 1))
    X[:n_lo, :d], X[n_lo:, :d], X[n_lo:, d] = x_lo, x_hi, 1.0
    ls = np.linspace(0.5, 1.0, d)
    theta = np.concatenate([[1.0], ls, [1.0], ls, [1.0]])
    return X, np.concatenate([y_lo, y_hi])[:, None], theta, 1e-3
