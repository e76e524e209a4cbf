"""GPU tests of the device-resident training loop (mfgp_gpr_batched_adam) and its host class
MultiBinMFGP against (a) the oracle's emulation of the reference loop (linear.py:190-221: TF Adam on unconstrained
variables, float32-rounded hypers, CosineDecay) and (b) the same loop driven from the host through the C-ABI."""
import numpy as np
import pytest

from oracle import mfgp_oracle as onp
from oracle import mfgp_oracle_torch as otc
from tests._helpers import dsoftplus

pytestmark = pytest.mark.gpu


def oracle_trajectory(X, y, theta0, noise, lr, n_steps, decay_steps=None, fix_rho=False):
    u = onp.softplus_inv(theta0).copy()
    opt = onp.TFAdam(lr=lr, cosine_decay_steps=decay_steps)
    hist = []
    for _ in range(n_steps):
        theta = onp.softplus(u)
        lml, g, _ = otc.gpr_lml_value_and_grad(X, y, theta, noise)
        hist.append(-lml)
        gu = -g * dsoftplus(theta)
        if fix_rho:
            gu[0] = 0.0
        opt.step([u], [gu])
    return np.array(hist), onp.softplus(u)


@pytest.mark.parametrize("cosine,fix_rho", [(False, False), (True, False), (False, True)])
def test_device_adam_matches_reference_loop(cosine, fix_rho):
    from multi_fidelity_gpflow_b200.multibin import MultiBinMFGP

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], np.ascontiguousarray(ds["Y"][:, [0, 17, 48]])
    steps, lr = 40, 0.05
    model = MultiBinMFGP(X, Y, num_restarts=2, seed=3, use_rho=not fix_rho)
    th0 = model.thetas.copy()
    model.optimize(max_iters=steps, learning_rate=lr, use_cosine_decay=cosine)
    assert model.loss_history.shape == (steps, 2, 3)
    for r in range(2):
        for p in range(3):
            hist, th_end = oracle_trajectory(X, Y[:, p:p + 1], th0[r, p], 1e-3, lr, steps, steps if cosine else None, fix_rho)
            # the loss before every update follows the reference trajectory (tolerance: gradient parity 1e-7 compounded)
            np.testing.assert_allclose(model.loss_history[:, r, p], hist, rtol=1e-7, atol=1e-7)
            np.testing.assert_allclose(model.thetas[r, p], th_end, rtol=1e-6)
            if fix_rho:
                assert model.thetas[r, p, 0] == th0[r, p, 0]


def test_split_runs_continue_the_same_trajectory_and_selection():
    from multi_fidelity_gpflow_b200.multibin import MultiBinMFGP

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"]
    a = MultiBinMFGP(X, Y, num_restarts=3, seed=1).optimize(max_iters=30, learning_rate=0.02)
    b = MultiBinMFGP(X, Y, num_restarts=3, seed=1)
    b.optimize(max_iters=10, learning_rate=0.02).optimize(max_iters=20, learning_rate=0.02)  # moments and step count carry over
    np.testing.assert_array_equal(a.loss_history, b.loss_history)
    np.testing.assert_array_equal(a.u, b.u)
    assert np.all(a.loss_history[-1] < a.loss_history[0])  # every (restart, bin) improved
    best = a.best_restart()
    loss = a.training_loss()
    assert np.array_equal(best, loss.argmin(axis=0))
    # prediction at the training inputs of the best model reproduces the HF data to the noise level
    hf = X[X[:, -1] == 1.0]
    mean, var = a.predict_f(hf)
    assert mean.shape == (hf.shape[0], 49) and np.all(var > 0)
    assert np.abs(mean - Y[X[:, -1] == 1.0]).max() < 0.2


def test_device_loop_equals_host_driven_loop_bitwise():
    """The same arithmetic driven step by step from the host (softplus / chain rule / Adam in NumPy around
    mfgp_gpr_batched_nlml_grad) agrees to rounding with the device loop."""
    from multi_fidelity_gpflow_b200 import _lib
    from multi_fidelity_gpflow_b200.multibin import MultiBinMFGP

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], np.ascontiguousarray(ds["Y"][:, :8])
    model = MultiBinMFGP(X, Y, num_restarts=2, seed=7)
    u = model.u.copy()
    model.optimize(max_iters=15, learning_rate=0.03)
    h = _lib.default_handle()
    opt = onp.TFAdam(lr=0.03)
    B = u.shape[0]
    hist = []
    for _ in range(15):
        th = onp.softplus(u)
        nlml, g = h.gpr_batched_nlml_grad(X, Y, th, np.full(B, 1e-3))
        hist.append(nlml.copy())
        opt.step([u], [g[:, :-1] * dsoftplus(th)])
    np.testing.assert_allclose(model.loss_history.reshape(15, B), np.array(hist), rtol=1e-12)
    np.testing.assert_allclose(model.u, u, rtol=1e-10, atol=1e-12)
