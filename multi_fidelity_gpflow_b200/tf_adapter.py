"""Optional TensorFlow glue: the reference's own `tf.GradientTape` training loops (mfgpflow/linear.py:203-209,
singlebin_svgp.py:79-85, linear_svgp.py:181-190) run UNCHANGED on top of libmfgp.so.

    from multi_fidelity_gpflow_b200.tf_adapter import TFMultiFidelityGPR
    model = TFMultiFidelityGPR(X, Y, d)                     # tf.Variables in GPflow's order, unconstrained (softplus)
    with tf.GradientTape() as tape:
        loss = -model.log_marginal_likelihood()             # one mfgp_gpr_nlml_grad call (value AND gradient)
    grads = tape.gradient(loss, model.trainable_variables)  # analytic gradient, chain rule 1 - exp(-theta)
    optimizer.apply_gradients(zip(grads, model.trainable_variables))

How: the objective is a `tf.custom_gradient` whose forward runs the CUDA path through `tf.py_function` and keeps the
analytic gradient for the backward function, so TensorFlow never differentiates through the kernel (the reference spends
most of its step in the scatter_nd backward).  Eager tensors that live on the GPU are handed over as
DLPack capsules (`tf.experimental.dlpack.to_dlpack` -> `_lib._ptr`), zero-copy.

TensorFlow / GPflow are NOT installable in this image, so this module is import-guarded and NOT exercised by
the test-suite: the plumbing it wraps -- unconstrained variables in `trainable_variables` order, softplus chain rule,
Keras-Adam with float32 hypers -- is the one `linear.py` / `optimizers.py` implement and tests/test_models_gpu.py checks
against the reference's recorded 500-step Adam trajectory (golden G3).  The DLPack consumer is tested with NumPy / torch
exporters (tests/test_abi.py, tests/test_cuda_kernels.py).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _tf():
    try:
        import tensorflow as tf
    except ImportError as e:  # pragma: no cover - TensorFlow is absent from this image
        raise ImportError("multi_fidelity_gpflow_b200.tf_adapter needs TensorFlow (the reference pins tensorflow~=2.10)") from e
    return tf


def to_library_buffer(t):
    """A TF tensor as something `_lib._ptr` accepts: GPU-resident eager tensors as DLPack capsules (zero-copy), anything
    else as a float64 NumPy array."""
    tf = _tf()
    if isinstance(t, tf.Tensor) and "GPU" in (t.device or "") and t.dtype == tf.float64:
        return tf.experimental.dlpack.to_dlpack(t)
    return np.ascontiguousarray(t.numpy() if hasattr(t, "numpy") else t, dtype=np.float64)


def softplus_inverse(x):
    x = np.asarray(x, dtype=np.float64)
    return x + np.log(-np.expm1(-x))


class TFMultiFidelityGPR:
    """tf.Variable-backed stand-in for the reference's MultiFidelityGPModel (linear.py:138-234) for GradientTape loops:
    `trainable_variables` = [rho (P, 1), kernel_L.lengthscales (d,), kernel_L.variance (), kernel_delta.lengthscales (d,),
    kernel_delta.variance ()] as UNCONSTRAINED float64 variables (gpflow.utilities.positive() = softplus); the Gaussian
    noise (1e-3) is fixed like linear.py:151-154; only rho[0] enters the kernel (quirk Q2)."""

    def __init__(self, X, Y, handle=None, noise=1e-3):
        tf = _tf()
        self._X = np.ascontiguousarray(X, dtype=np.float64)
        self._Y = np.ascontiguousarray(Y, dtype=np.float64)
        self._d = self._X.shape[1] - 1
        self._P = self._Y.shape[1]
        self._h = handle or _lib.default_handle()
        self.noise = float(noise)
        one = lambda shape: tf.Variable(softplus_inverse(np.ones(shape)), dtype=tf.float64)
        self.rho, self.ls_L, self.var_L, self.ls_delta, self.var_delta = one((self._P, 1)), one((self._d,)), one(()), one((self._d,)), one(())
        self.trainable_variables = [self.rho, self.ls_L, self.var_L, self.ls_delta, self.var_delta]

    def _objective(self):
        tf = _tf()
        d, P = self._d, self._P

        @tf.custom_gradient
        def nlml(rho_u, lsL_u, vL_u, lsD_u, vD_u):
            def forward(rho_u, lsL_u, vL_u, lsD_u, vD_u):
                sp = lambda u: np.logaddexp(0.0, u.numpy())
                theta = np.concatenate([[sp(rho_u)[0, 0]], sp(lsL_u), [sp(vL_u)], sp(lsD_u), [sp(vD_u)]])
                try:
                    val, g = self._h.gpr_nlml_grad(self._X, self._Y, theta, self.noise)
                except _lib.NotPositiveDefiniteError as e:  # the error the reference's loop would see from tf.linalg.cholesky
                    raise tf.errors.InvalidArgumentError(None, None, f"Cholesky decomposition was not successful. {e}")
                gu = g[: 2 * d + 3] * (1.0 - np.exp(-theta))  # d theta / d u = sigmoid(u) = 1 - exp(-theta)
                g_rho = np.zeros((P, 1))
                g_rho[0, 0] = gu[0]
                return [np.float64(val), g_rho, gu[1:1 + d], np.float64(gu[1 + d]), gu[2 + d:2 + 2 * d], np.float64(gu[2 + 2 * d])]

            out = tf.py_function(forward, [rho_u, lsL_u, vL_u, lsD_u, vD_u], [tf.float64] * 6)
            val, grads = out[0], out[1:]
            val.set_shape(())

            def backward(dy):
                return [dy * tf.reshape(g, tf.shape(v)) for g, v in zip(grads, (rho_u, lsL_u, vL_u, lsD_u, vD_u))]

            return val, backward

        return nlml(*self.trainable_variables)

    def training_loss(self):
        return self._objective()

    def log_marginal_likelihood(self):
        return -self._objective()
