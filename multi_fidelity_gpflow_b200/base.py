"""Host-side parameter containers mirroring the slice of gpflow.Parameter / gpflow.utilities the
reference uses (linear.py:5,46-52; singlebin_svgp.py:95; linear_svgp.py:6-7,112-115)."""
from __future__ import annotations

import numpy as np


class Transform:
    """theta = lower + softplus(u) (gpflow.utilities.positive(lower)), sigmoid(u) (tfp.bijectors.Sigmoid) or identity."""

    def __init__(self, kind="identity", lower=0.0):
        self.kind, self.lower = kind, float(lower)

    def forward(self, u):
        if self.kind == "identity":
            return np.array(u, dtype=np.float64, copy=True)
        if self.kind == "sigmoid":
            return 1.0 / (1.0 + np.exp(-np.asarray(u, dtype=np.float64)))
        return self.lower + np.logaddexp(0.0, u)

    def inverse(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        if self.kind == "identity":
            return theta.copy()
        if self.kind == "sigmoid":
            return np.log(theta) - np.log1p(-theta)
        x = theta - self.lower
        return x + np.log(-np.expm1(-x))

    def dtheta_du(self, theta):
        if self.kind == "identity":
            return np.ones_like(np.asarray(theta, dtype=np.float64))
        if self.kind == "sigmoid":
            t = np.asarray(theta, dtype=np.float64)
            return t * (1.0 - t)
        return 1.0 - np.exp(-(np.asarray(theta, dtype=np.float64) - self.lower))


def positive(lower=0.0):
    return Transform("softplus", lower)


def sigmoid():
    return Transform("sigmoid")


class Parameter:
    """Value stored UNCONSTRAINED (as a tf.Variable would be); `.numpy()` is the constrained value."""

    def __init__(self, value, transform=None, trainable=True, name=None):
        if isinstance(value, Parameter):
            transform = transform or value.transform
            value = value.numpy()
        self.transform = transform or Transform()
        self.unconstrained = self.transform.inverse(np.asarray(value, dtype=np.float64))
        self.trainable = bool(trainable)
        self.name = name

    def numpy(self):
        return self.transform.forward(self.unconstrained)

    def assign(self, value):
        self.unconstrained = self.transform.inverse(np.asarray(value, dtype=np.float64)).reshape(self.unconstrained.shape)

    @property
    def shape(self):
        return self.unconstrained.shape

    def grad_to_unconstrained(self, g_constrained):
        return np.asarray(g_constrained, dtype=np.float64).reshape(self.shape) * self.transform.dtheta_du(self.numpy())

    def __repr__(self):
        return f"Parameter(name={self.name}, shape={self.shape}, trainable={self.trainable}, value={self.numpy()})"


def set_trainable(obj, flag: bool):
    """gpflow.utilities.set_trainable for a Parameter or any object holding Parameters."""
    if isinstance(obj, Parameter):
        obj.trainable = bool(flag)
        return
    for p in parameters_of(obj).values():
        p.trainable = bool(flag)


def parameters_of(obj, prefix=""):
    """Ordered {dotted_name: Parameter} following the declared order `_param_order` of each object."""
    out = {}
    order = getattr(obj, "_param_order", None)
    if order is None:
        return out
    for attr in order:
        v = getattr(obj, attr)
        name = f"{prefix}.{attr}"
        if isinstance(v, Parameter):
            out[name] = v
        elif isinstance(v, (list, tuple)):
            for i, item in enumerate(v):
                out.update(parameters_of(item, f"{name}[{i}]"))
        elif v is not None:
            out.update(parameters_of(v, name))
    return out


def parameter_dict(model):
    """gpflow.utilities.parameter_dict: {'.kernel.kernels[0].kernel_L.lengthscales': value, ...} (constrained)."""
    return {k: p.numpy() for k, p in parameters_of(model).items()}


def multiple_assign(model, values):
    params = parameters_of(model)
    for k, v in values.items():
        params[k].assign(v)
