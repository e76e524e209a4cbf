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


def test_failed_restart_is_per_problem_status_not_an_abort():
    """with a per-problem info[] a non-positive pivot is that problem's status; the others train on."""
    from multi_fidelity_gpflow_b200 import _lib
    from multi_fidelity_gpflow_b200.optimizers import adam_step_factors

    h = _lib.default_handle()
    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], np.ascontiguousarray(ds["Y"][:, :4])
    B, steps = 4, 6
    u0 = onp.softplus_inv(np.ones((B, 13)))
    lr_t, b1, b2 = adam_step_factors(0.05, steps)
    good = np.full(B, 1e-3)
    bad = good.copy()
    bad[2] = -5.0  # K - 5 I: not positive definite from the first pivot on
    ref_u, ref_hist = u0.copy(), np.empty((steps, B))
    h.gpr_batched_adam(X, Y, ref_u, np.zeros_like(u0), np.zeros_like(u0), good, lr_t, b1, b2, loss_hist=ref_hist)
    u, hist, info = u0.copy(), np.empty((steps, B)), np.zeros(B, dtype=np.int32)
    h.gpr_batched_adam(X, Y, u, np.zeros_like(u0), np.zeros_like(u0), bad, lr_t, b1, b2, loss_hist=hist, info=info)  # no raise
    assert info[2] > 0 and np.all(info[[0, 1, 3]] == 0)
    keep = [0, 1, 3]
    np.testing.assert_array_equal(u[keep], ref_u[keep])
    np.testing.assert_array_equal(hist[:, keep], ref_hist[:, keep])
    assert not np.isfinite(hist[:, 2]).any()
    with pytest.raises(_lib.NotPositiveDefiniteError):  # without info[] the status is the call's error, as before
        h.gpr_batched_adam(X, Y, u0.copy(), np.zeros_like(u0), np.zeros_like(u0), bad, lr_t, b1, b2)
    # single evaluation: same convention
    info2 = np.zeros(B, dtype=np.int32)
    nl, _ = h.gpr_batched_nlml_grad(X, Y, np.ones((B, 13)), bad, info=info2)
    assert info2[2] > 0 and np.isfinite(nl[keep]).all()
    with pytest.raises(_lib.NotPositiveDefiniteError):
        h.gpr_batched_nlml_grad(X, Y, np.ones((B, 13)), bad)
