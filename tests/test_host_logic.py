"""CPU tests of the host-side mirror: transforms, TF-Adam / CosineDecay emulation, SciPy packing,
GPflow-style parameter names, W initialisers, fixture loader, multi-process sharding (gloo)."""
import os
import sys

import numpy as np
import pytest

from multi_fidelity_gpflow_b200 import base, optimizers
from multi_fidelity_gpflow_b200.data import PowerSpecs
from multi_fidelity_gpflow_b200.dist import shard_range
from oracle import mfgp_oracle as onp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_softplus_transform_roundtrip_and_chain_rule():
    for lower in (0.0, 1e-6):
        t = base.positive(lower)
        theta = np.array([1e-3, 0.1, 1.0, 30.0]) + lower
        u = t.inverse(theta)
        np.testing.assert_allclose(t.forward(u), theta, rtol=1e-13)
        eps = 1e-6
        fd = (t.forward(u + eps) - t.forward(u - eps)) / (2 * eps)
        np.testing.assert_allclose(t.dtheta_du(theta), fd, rtol=1e-6)
    p = base.Parameter(1e-3, transform=base.positive(1e-6))
    assert abs(p.numpy() - 1e-3) < 1e-18 + 1e-15


def test_adam_matches_oracle_tf_adam():
    rng = np.random.default_rng(0)
    for decay in (None, 50):
        p1 = base.Parameter(rng.standard_normal((3, 2)))
        p2 = base.Parameter(rng.standard_normal(4))
        ref = [p1.unconstrained.copy(), p2.unconstrained.copy()]
        lr = optimizers.CosineDecay(0.1, decay) if decay else 0.1
        opt = optimizers.Adam(lr)
        oref = onp.TFAdam(lr=0.1, cosine_decay_steps=decay)
        for step in range(60):
            g = [rng.standard_normal((3, 2)), rng.standard_normal(4)]
            g[1][2] = 0.0  # zero gradient => exactly zero update (quirk Q5 relies on it)
            opt.apply_gradients(zip([x.copy() for x in g], [p1, p2]))
            oref.step(ref, g)
            assert np.array_equal(p1.unconstrained, ref[0]) and np.array_equal(p2.unconstrained, ref[1])


def test_adam_float32_hypers():
    opt = optimizers.Adam(0.1)
    assert opt.lr == float(np.float32(0.1)) and opt.b1 == float(np.float32(0.9)) and opt.b2 == float(np.float32(0.999))
    assert opt.lr != 0.1


def test_scipy_packs_unconstrained_in_order():
    a = base.Parameter(np.array([1.0, 2.0]), transform=base.positive())
    b = base.Parameter(np.array([[0.5]]))
    target = np.array([0.3, -0.2, 1.5])

    def vg():
        x = np.concatenate([a.unconstrained.ravel(), b.unconstrained.ravel()])
        return 0.5 * np.sum((x - target) ** 2), [(x - target)[:2], (x - target)[2:].reshape(1, 1)]

    res = optimizers.Scipy().minimize(vg, [a, b], options={"maxiter": 100})
    assert res.success
    np.testing.assert_allclose(np.concatenate([a.unconstrained, b.unconstrained.ravel()]), target, atol=1e-6)


def test_parameter_dict_names_match_gpflow_layout():
    from multi_fidelity_gpflow_b200.kernels import SeparateIndependent, SquaredExponential, replicate_mf_kernels
    from multi_fidelity_gpflow_b200.likelihoods import Gaussian
    from multi_fidelity_gpflow_b200.svgp_base import SVGPBase

    m = SVGPBase()
    ks = replicate_mf_kernels(SquaredExponential(lengthscales=np.ones(3)), SquaredExponential(lengthscales=np.ones(3)), 2)
    m._init_svgp(SeparateIndependent(ks), Gaussian(), np.zeros((4, 4)), 2)
    names = list(base.parameter_dict(m))
    assert ".kernel.kernels[0].kernel_L.lengthscales" in names and ".kernel.kernels[1].rho" in names
    assert ".likelihood.variance" in names and ".inducing_variable.inducing_variable.Z" in names
    assert ".q_mu" in names and ".q_sqrt" in names
    shapes = [p.shape for p in m.trainable_variables]
    # order printed by the reference notebook: q_mu, q_sqrt, Z, then (rho, ls, var, ls, var) per kernel, likelihood
    assert shapes[:3] == [(4, 2), (2, 4, 4), (4, 4)] and shapes[3:8] == [(1, 1), (3,), (), (3,), ()] and shapes[-1] == ()
    # deep-copied base kernels: independent parameters per output (singlebin_svgp.py:39)
    ks[0].kernel_L.variance.assign(2.0)
    assert ks[1].kernel_L.variance.numpy() == 1.0


def test_power_specs_loader_matches_oracle_loader():
    for name in ("hbs", "goku"):
        ps = PowerSpecs().read_from_npz(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
        X, Y = ps.training_arrays()
        ds = onp.load_dataset(name)
        assert np.array_equal(X, ds["X"]) and np.array_equal(Y, ds["Y"])
        Xt, Yt = ps.test_arrays()
        assert np.array_equal(Xt, ds["X_test"]) and np.array_equal(Yt, ds["Y_test"])


def test_shard_range_partitions():
    for n in (0, 1, 7, 49, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from multi_fidelity_gpflow_b200.dist import dp_svgp_value_and_grad, gather_bins, shard_range
    from oracle import mfgp_oracle as onp
    from oracle import mfgp_oracle_torch as otc

    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"][:, :6]
    rng = np.random.default_rng(0)
    # (1) bins shard with no data-path collective; gather only assembles the answer
    B = Y.shape[1]
    th = np.tile(onp.default_theta(5), (B, 1)) * np.exp(0.1 * rng.standard_normal((B, 13)))
    nz = np.full(B, 1e-3)
    lo, hi = shard_range(B, rank, world)
    vals, grads = otc.gpr_batched_value_and_grad(X, Y[:, lo:hi], th[lo:hi], nz[lo:hi])
    full_v = gather_bins(vals, B)
    full_g = gather_bins(grads, B)
    # (2) data-parallel SVGP: rows shard, one all-reduce
    M, L, P = 8, 3, 6
    Z = X[rng.permutation(53)[:M]].copy()
    ths = np.tile(onp.default_theta(5), (L, 1))
    W = onp.initialize_W(P, L, 0.4, 0.2)
    q_mu, q_sqrt = 0.1 * rng.standard_normal((M, L)), np.tile(0.3 * np.eye(M), (L, 1, 1))

    def local_fn(Xr, Yr, scale, klm):
        return otc.svgp_value_and_grad(Xr, Yr, Z, ths, q_mu, q_sqrt, 0.8, W, num_data=scale * Xr.shape[0], kl_mult=klm)

    out = dp_svgp_value_and_grad(local_fn, X, Y, num_data=53, kl_mult=1.7)
    # (3) the data-parallel TRAINING loop (dist.dp_svgp_adam): flat buffers, in-place all-reduce, Adam on every rank;
    #     block arithmetic by the oracle double, schedule / sharding / reduce convention by the product code
    from multi_fidelity_gpflow_b200.dist import dp_svgp_adam
    from multi_fidelity_gpflow_b200.optimizers import adam_step_factors
    from tests._helpers import OracleSvgpOps

    solo = dist.new_group([0])
    u0, mask, shape = _dp_loop_problem(onp, X, Y, Z, ths, W, q_mu, q_sqrt)
    lr_t, b1, b2 = adam_step_factors(0.05, 3, cosine_decay_steps=3)
    res2 = dp_svgp_adam(OracleSvgpOps(), X, Y, shape, u0, mask, lr_t, b1, b2, num_data=53, kl_mult=1.7)
    res1 = dp_svgp_adam(OracleSvgpOps(), X, Y, shape, u0, mask, lr_t, b1, b2, num_data=53, kl_mult=1.7, group=solo) if rank == 0 else None
    if rank == 0:
        q.put((full_v, full_g, out, res2, res1))
    else:
        q.put(("rank1", res2))
    dist.destroy_process_group()


def _dp_loop_problem(onp, X, Y, Z, ths, W, q_mu, q_sqrt):
    L, M, P, d = ths.shape[0], Z.shape[0], W.shape[0], X.shape[1] - 1
    u0 = np.concatenate([onp.softplus_inv(ths).ravel(), Z.ravel(), W.ravel(), q_mu.ravel(), np.tril(q_sqrt).ravel(),
                         onp.softplus_inv(np.array([0.8 - 1e-6]))])
    mask = np.ones(u0.size, dtype=np.uint8)
    mask[L * (2 * d + 3) + d::d + 1][:M] = 1  # (fidelity column trainable in name; its gradient is exactly zero, quirk Q5)
    mask[0] = 0  # rho of latent 0 frozen: a masked entry must not move on any rank
    shape = dict(L=L, M=M, P=P, d=d, has_W=True, hetero=False, masked=False, lik_per_output=False, lik_lower=1e-6)
    return u0, mask, shape


def test_two_rank_gloo_sharding_matches_single_process():
    import torch.multiprocessing as mp

    from oracle import mfgp_oracle_torch as otc

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=480), q.get(timeout=480)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (full_v, full_g, out, res2, res1), = [g for g in got if len(g) == 5]
    (_, res2_rank1), = [g for g in got if len(g) == 2]
    # data-parallel loop: both ranks end bit-identical; two ranks follow the one-rank trajectory to rounding
    for a, b in zip(res2, res2_rank1):
        assert np.array_equal(a, b)
    np.testing.assert_allclose(res2[0], res1[0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res2[1], res1[1], rtol=1e-11)
    np.testing.assert_allclose(res2[2], res1[2], rtol=1e-11)
    assert res2[1][-1] < res2[1][0] and res2[0][0] == res1[0][0]  # the loss went down; the frozen entry never moved
    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"][:, :6]
    rng = np.random.default_rng(0)
    th = np.tile(onp.default_theta(5), (6, 1)) * np.exp(0.1 * rng.standard_normal((6, 13)))
    v, g = otc.gpr_batched_value_and_grad(X, Y, th, np.full(6, 1e-3))
    assert np.array_equal(full_v, v) and np.array_equal(full_g, g)  # no arithmetic crosses ranks: bit-identical
    M, L, P = 8, 3, 6
    Z = X[rng.permutation(53)[:M]].copy()
    ths = np.tile(onp.default_theta(5), (L, 1))
    W = onp.initialize_W(P, L, 0.4, 0.2)
    q_mu, q_sqrt = 0.1 * rng.standard_normal((M, L)), np.tile(0.3 * np.eye(M), (L, 1, 1))
    ref = otc.svgp_value_and_grad(X, Y, Z, ths, q_mu, q_sqrt, 0.8, W, num_data=53, kl_mult=1.7)
    assert abs(out["elbo"] - ref["elbo"]) < 1e-11 * abs(ref["elbo"])
    for k in ("g_Z", "g_thetas", "g_q_mu", "g_q_sqrt", "g_W"):
        np.testing.assert_allclose(out[k], ref[k], rtol=1e-10, atol=1e-12 * np.abs(ref[k]).max())
    assert abs(out["g_lik_var"] - ref["g_lik_var"]) < 1e-10 * abs(ref["g_lik_var"])


def test_adam_step_factors_match_tf_adam_emulation():
    """The per-step factors handed to the device-resident loops reproduce the oracle's TF-Adam (fp32 hypers, CosineDecay)."""
    from multi_fidelity_gpflow_b200.optimizers import adam_step_factors

    for lr, decay in ((0.01, None), (0.005, 40), (0.1, 7)):
        f, b1, b2 = adam_step_factors(lr, 25, cosine_decay_steps=decay)
        ref = onp.TFAdam(lr=lr, cosine_decay_steps=decay)
        assert b1 == ref.b1 and b2 == ref.b2
        u = np.array([0.3, -1.2])
        m, v = np.zeros(2), np.zeros(2)
        ur = u.copy()
        rng = np.random.default_rng(0)
        for s in range(25):
            g = rng.standard_normal(2)
            ref.step([ur], [g])
            m += (g - m) * (1.0 - b1)
            v += (g * g - v) * (1.0 - b2)
            u -= f[s] * m / (np.sqrt(v) + 1e-7)
            np.testing.assert_allclose(u, ur, rtol=1e-14)
    # resuming: factors of steps 10.. equal the tail of the full schedule (no decay; a decay restarts like the reference)
    full, _, _ = adam_step_factors(0.01, 30)
    tail, _, _ = adam_step_factors(0.01, 20, first_step=10)
    np.testing.assert_array_equal(full[10:], tail)


def test_svgp_flat_parameter_layout_matches_the_c_abi():
    """mfgp_svgp_adam's flat layout [theta | Z | W | q_mu | q_sqrt | lik_var]: sizes, order, freeze mask source."""
    from multi_fidelity_gpflow_b200.kernels import SquaredExponential
    from multi_fidelity_gpflow_b200.linear_svgp import LatentMFCoregionalizationSVGP

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"]
    d, L, M, P = 5, 3, 12, 49
    mdl = LatentMFCoregionalizationSVGP(X, Y, SquaredExponential(lengthscales=np.ones(d)), SquaredExponential(lengthscales=np.ones(d)),
                                        num_latents=L, num_inducing=M, num_outputs=P)
    items, n = mdl._flat_parameters(d)
    assert n == L * (2 * d + 3) + M * (d + 1) + P * L + M * L + L * M * M + 1
    sizes = [sl.stop - sl.start for _, sl in items]
    assert sizes[:5] == [1, d, 1, d, 1] and sizes[-5:] == [M * (d + 1), P * L, M * L, L * M * M, 1]
    assert [sl.start for _, sl in items] == list(np.cumsum([0] + sizes[:-1]))
    assert items[-1][0] is mdl.likelihood.variance and items[-2][0] is mdl.q_sqrt
    shared = LatentMFCoregionalizationSVGP(X, Y, SquaredExponential(), SquaredExponential(), num_latents=L, num_inducing=M,
                                           num_outputs=P)
    with pytest.raises(NotImplementedError):
        shared._flat_parameters(d)  # one shared lengthscale: the device loop refuses instead of silently training d copies


class _FakeHandle:
    """Stand-in for _lib.Handle in HOST-LOGIC tests only (records calls, returns synthetic numbers; no arithmetic claims)."""

    def __init__(self):
        self.calls = []
        self.fail_problem = None  # flat index of a problem whose Cholesky "fails" (per-problem info, no exception)

    def gpr_batched_nlml_grad(self, X, Y, thetas, noises, want_grad=True, info=None, **kw):
        self.calls.append(("eval", thetas.shape, Y.shape))
        nlml = thetas[:, 0] * 10.0 + np.arange(thetas.shape[0]) % Y.shape[1]  # depends on rho and on the bin
        if self.fail_problem is not None and info is not None:
            info[self.fail_problem] = 3
        return nlml, None

    def gpr_batched_adam(self, X, Y, u, m, v, noises, lr_t, b1, b2, eps, fix_rho=False, loss_hist=None, theta_out=None, info=None):
        self.calls.append(("adam", u.shape, len(lr_t), fix_rho))
        u -= 0.01 * len(lr_t)  # "training": every unconstrained variable moves the same way
        loss_hist[:] = np.arange(len(lr_t))[:, None] + np.arange(u.shape[0])[None, :]
        if self.fail_problem is not None:
            info[self.fail_problem] = 3
            loss_hist[:, self.fail_problem] = np.nan
        return loss_hist, theta_out

    def gpr_predict(self, X, y, Xnew, theta, noise):
        self.calls.append(("predict", y.shape, tuple(theta[:1])))
        return np.full((Xnew.shape[0], 1), theta[0]), np.full(Xnew.shape[0], noise)


def test_multibin_host_logic_with_fake_handle():
    """MultiBinMFGP: restart 0 starts at the reference's initial values, problem b = r * P + p, moments / step count carry over,
    best-restart selection ignores non-finite losses, predictions use each bin's best hyper-parameters."""
    from multi_fidelity_gpflow_b200.multibin import MultiBinMFGP

    rng = np.random.default_rng(0)
    X = np.hstack([rng.random((20, 3)), (np.arange(20) >= 15).astype(float)[:, None]])
    Y = rng.standard_normal((20, 4))
    fh = _FakeHandle()
    mdl = MultiBinMFGP(X, Y, num_restarts=3, seed=5, handle=fh)
    th = mdl.thetas
    assert th.shape == (3, 4, 9)
    np.testing.assert_allclose(th[0], 1.0, rtol=1e-14)           # restart 0: rho = lengthscales = variances = 1
    assert np.all(th[1:] > 0) and not np.allclose(th[1], th[2])   # perturbed restarts, all positive
    assert np.array_equal(MultiBinMFGP(X, Y, num_restarts=3, seed=5, handle=fh).u, mdl.u)  # seeded
    u0 = mdl.u.copy()
    mdl.optimize(max_iters=7, learning_rate=0.1)
    mdl.optimize(max_iters=5, learning_rate=0.1, use_cosine_decay=True)
    assert mdl.iterations == 12 and mdl.loss_history.shape == (12, 3, 4)
    assert [c for c in fh.calls if c[0] == "adam"] == [("adam", (12, 9), 7, False), ("adam", (12, 9), 5, False)]
    np.testing.assert_allclose(mdl.u, u0 - 0.12)
    # loss_history[s, r, p] is problem b = r * P + p of the flat batch
    np.testing.assert_array_equal(mdl.loss_history[:7, 2, 1], np.arange(7) + (2 * 4 + 1))
    # selection: lowest NLML per bin; the fake loss grows with rho, so the restart with the smallest rho wins every bin
    best = mdl.best_restart()
    assert np.array_equal(best, np.argmin(mdl.thetas[:, :, 0] * 10.0, axis=0))
    mean, var = mdl.predict_f(X[:5])
    assert mean.shape == (5, 4) and var.shape == (5, 4)
    np.testing.assert_allclose(mean[0], mdl.best_thetas()[:, 0])  # each bin predicted with its own best theta
    # a restart whose Cholesky fails is recorded, not raised: counters advance, selection skips it
    fh.fail_problem = int(best[1]) * 4 + 1  # the restart that is currently best for bin 1
    mdl.optimize(max_iters=3, learning_rate=0.1)
    assert mdl.iterations == 15 and mdl.loss_history.shape == (15, 3, 4)
    assert mdl.failed[best[1], 1] == 3 and np.count_nonzero(mdl.failed) == 1
    assert np.isnan(mdl.training_loss()[best[1], 1])
    assert mdl.best_restart()[1] != best[1]
    fh.fail_problem = None
    frozen = MultiBinMFGP(X, Y, num_restarts=1, use_rho=False, handle=fh)
    frozen.optimize(max_iters=2)
    assert fh.calls[-1] == ("adam", (4, 9), 2, True)
    with pytest.raises(ValueError):
        MultiBinMFGP(np.zeros((65, 3)), np.zeros((65, 1)), handle=fh)  # the small-matrix kernel is N <= 64


def test_synthetic_two_fidelity_is_the_survey_config_c5():
    """data.synthetic_two_fidelity (what bench.py's exact-GP leg feeds the product) is SURVEY 8(d) config C5 -- the same
    arrays as the oracle's restatement of that recipe, so bench and parity tests talk about the same problem."""
    from multi_fidelity_gpflow_b200.data import synthetic_two_fidelity

    for N in (64, 1000):
        X, Y, theta, noise = synthetic_two_fidelity(N)
        ds = onp.synthetic_exact_dataset(N)
        assert np.array_equal(X, ds["X"]) and np.array_equal(Y, ds["Y"]) and np.array_equal(theta, ds["theta"]) and noise == ds["noise"]
        assert X.shape == (N, 11) and int((X[:, -1] == 1).sum()) == N // 8
        hf = X[X[:, -1] == 1, :-1]
        lf = X[X[:, -1] == 0, :-1]
        assert all(any(np.array_equal(h, l) for l in lf) for h in hf[:5])  # nested design: HF inputs are LF inputs
