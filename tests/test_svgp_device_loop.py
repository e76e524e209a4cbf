"""GPU tests: the device-resident SVGP training loop (mfgp_svgp_adam) follows the same trajectory as the models' host
optimize() loops (mirrors of mfgpflow/singlebin_svgp.py:64-97 and mfgpflow/linear_svgp.py:153-203)."""
import copy

import numpy as np
import pytest

from oracle import mfgp_oracle as onp

pytestmark = pytest.mark.gpu


def _kernels(d):
    from multi_fidelity_gpflow_b200.kernels import SquaredExponential

    return SquaredExponential(lengthscales=np.ones(d)), SquaredExponential(lengthscales=np.ones(d))


def _params(model):
    out = [model.q_mu.numpy(), np.tril(model.q_sqrt.numpy()), model.Z.numpy(), np.ravel(model.likelihood.variance.numpy())]
    W = getattr(model.kernel, "W", None)
    if W is not None:
        out.append(W.numpy())
    out.append(np.stack([k.theta(5) for k in model.kernel.kernels]))
    return out


def test_singlebin_svgp_device_loop_equals_host_loop():
    from multi_fidelity_gpflow_b200.singlebin_svgp import SingleBinSVGP

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], np.ascontiguousarray(ds["Y"][:, :6])
    kL, kD = _kernels(5)
    a = SingleBinSVGP(X, Y, kL, kD, 6, ds["Z_kmeans50"])
    b = copy.deepcopy(a)
    a.optimize((X, Y), max_iters=12, initial_lr=0.01, verbose=False)
    b.optimize_on_device((X, Y), max_iters=12, initial_lr=0.01)
    np.testing.assert_allclose(b.loss_history, a.loss_history, rtol=1e-10)
    for pa, pb in zip(_params(a), _params(b)):
        np.testing.assert_allclose(pb, pa, rtol=1e-8, atol=1e-10)
    # fidelity column of Z never moves (quirk Q5: its gradient is exactly zero)
    assert np.array_equal(b.Z.numpy()[:, -1], ds["Z_kmeans50"][:, -1]) or np.allclose(b.Z.numpy()[:, -1], a.Z.numpy()[:, -1], rtol=0, atol=0)


@pytest.mark.parametrize("klm", [1.0, 2.5])
def test_latent_svgp_device_loop_equals_host_loop(klm):
    from multi_fidelity_gpflow_b200.linear_svgp import LatentMFCoregionalizationSVGP

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"]
    kL, kD = _kernels(5)
    a = LatentMFCoregionalizationSVGP(X, Y, kL, kD, num_latents=4, num_inducing=20, num_outputs=49)
    b = copy.deepcopy(a)
    a.optimize((X, Y), max_iters=10, initial_lr=0.005, kl_multiplier=klm, verbose=False)
    b.optimize_on_device((X, Y), max_iters=10, initial_lr=0.005, kl_multiplier=klm)
    np.testing.assert_allclose(b.loss_history, a.loss_history, rtol=1e-10)
    np.testing.assert_allclose(b.kl_history, a.kl_history, rtol=1e-10, atol=1e-12)
    for pa, pb in zip(_params(a), _params(b)):
        np.testing.assert_allclose(pb, pa, rtol=1e-8, atol=1e-10)


def test_frozen_parameters_stay_put():
    from multi_fidelity_gpflow_b200.base import set_trainable
    from multi_fidelity_gpflow_b200.linear_svgp import LatentMFCoregionalizationSVGP

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"]
    kL, kD = _kernels(5)
    mdl = LatentMFCoregionalizationSVGP(X, Y, kL, kD, num_latents=3, num_inducing=16, num_outputs=49, w_type="fixed_independent")
    set_trainable(mdl.likelihood.variance, False)
    W0, lv0 = mdl.kernel.W.numpy().copy(), mdl.likelihood.variance.numpy().copy()
    mdl.optimize_on_device((X, Y), max_iters=5, initial_lr=0.01)
    assert np.array_equal(mdl.kernel.W.numpy(), W0) and np.array_equal(mdl.likelihood.variance.numpy(), lv0)
    assert mdl.loss_history[-1] < mdl.loss_history[0]


@pytest.mark.parametrize("kind", ["hetero", "masked"])
def test_likelihood_variants_device_loop_equals_host_loop(kind):
    """HeteroscedasticGaussian's variance transform has lower bound 0 (linear_svgp.py:240), not the 1e-6 of
    gpflow's Gaussian; MaskedGaussian trains one variance per output.  Both loops must follow the same trajectory."""
    from multi_fidelity_gpflow_b200.linear_svgp import LatentMFCoregionalizationSVGP

    ds = onp.load_dataset("hbs")
    rng = np.random.default_rng(5)
    X, Y = ds["X"], ds["Y"].copy()
    kL, kD = _kernels(5)
    if kind == "hetero":
        Yt = np.hstack([Y, 0.05 + 0.1 * rng.random(Y.shape)])
        a = LatentMFCoregionalizationSVGP(X, Yt, kL, kD, num_latents=4, num_inducing=20, heterosed=True)
        a.likelihood.variance.assign(np.array([2e-6]))  # close to 0: a 1e-6 lower bound would change theta by 50 %
        assert a.likelihood.variance.transform.lower == 0.0
    else:
        Y[rng.random(Y.shape) < 0.25] = np.nan
        Yt = Y
        a = LatentMFCoregionalizationSVGP(X, Yt, kL, kD, num_latents=4, num_inducing=20, num_outputs=49, masked=True)
        assert a.likelihood.variance.shape == (49,)
    b = copy.deepcopy(a)
    a.optimize((X, Yt), max_iters=8, initial_lr=0.01, verbose=False)
    b.optimize_on_device((X, Yt), max_iters=8, initial_lr=0.01)
    np.testing.assert_allclose(b.loss_history, a.loss_history, rtol=1e-10)
    for pa, pb in zip(_params(a), _params(b)):
        np.testing.assert_allclose(pb, pa, rtol=1e-8, atol=1e-10)


def test_device_loop_history_semantics_follow_the_reference_loops():
    """linear_svgp.py:194 runs range(len(loss_history), max_iters) with a fresh optimizer; singlebin_svgp.py:79 resets."""
    from multi_fidelity_gpflow_b200.linear_svgp import LatentMFCoregionalizationSVGP
    from multi_fidelity_gpflow_b200.singlebin_svgp import SingleBinSVGP

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"]
    kL, kD = _kernels(5)
    a = LatentMFCoregionalizationSVGP(X, Y, kL, kD, num_latents=3, num_inducing=16, num_outputs=49)
    b = copy.deepcopy(a)
    a.optimize((X, Y), max_iters=4, initial_lr=0.01, verbose=False)
    a.optimize((X, Y), max_iters=9, initial_lr=0.01, verbose=False)  # 5 more steps, optimizer restarted
    b.optimize_on_device((X, Y), max_iters=4, initial_lr=0.01)
    b.optimize_on_device((X, Y), max_iters=9, initial_lr=0.01)
    assert len(b.loss_history) == len(a.loss_history) == 9 and len(b.kl_history) == 9
    np.testing.assert_allclose(b.loss_history, a.loss_history, rtol=1e-10)
    b.optimize_on_device((X, Y), max_iters=9, initial_lr=0.01)  # nothing left to do
    assert len(b.loss_history) == 9
    Y6 = np.ascontiguousarray(Y[:, :6])
    s = SingleBinSVGP(X, Y6, kL, kD, 6, ds["Z_kmeans50"])
    s.optimize_on_device((X, Y6), max_iters=3, initial_lr=0.01)
    s.optimize_on_device((X, Y6), max_iters=5, initial_lr=0.01)
    assert len(s.loss_history) == 5  # reset, like the reference


def test_data_parallel_loop_on_one_rank_equals_device_loop():
    """dist.dp_svgp_adam (mfgp_svgp_constrain / _elbo_grad_flat / _adam_update + an in-place NCCL all-reduce) with a
    one-rank NCCL group: the same trajectory as mfgp_svgp_adam.  Runs on the single-GPU test box; the two-rank version
    is tests/test_multi_gpu.py and the bench's --gpus N leg."""
    import os

    import torch
    import torch.distributed as dist

    from multi_fidelity_gpflow_b200.linear_svgp import LatentMFCoregionalizationSVGP

    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(29700 + os.getpid() % 200))
        torch.cuda.set_device(0)
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        ds = onp.load_dataset("hbs")
        X, Y = ds["X"], ds["Y"]
        kL, kD = _kernels(5)
        a = LatentMFCoregionalizationSVGP(X, Y, kL, kD, num_latents=4, num_inducing=20, num_outputs=49)
        b = copy.deepcopy(a)
        a.optimize_on_device((X, Y), max_iters=7, initial_lr=0.01, kl_multiplier=1.5)
        t = {}
        b.optimize_data_parallel((X, Y), max_iters=7, initial_lr=0.01, kl_multiplier=1.5, timing=t)
        assert t["ms_per_step"] > 0
        np.testing.assert_allclose(b.loss_history, a.loss_history, rtol=1e-12)
        np.testing.assert_allclose(b.kl_history, a.kl_history, rtol=1e-12, atol=1e-14)
        for pa, pb in zip(_params(a), _params(b)):
            np.testing.assert_allclose(pb, pa, rtol=1e-12, atol=1e-14)
    finally:
        if created:
            dist.destroy_process_group()
