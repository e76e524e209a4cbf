"""GPU parity and property tests of the batched small-GPR kernel K6 v4 (csrc/gpr_small_v4.cu) through the C-ABI:
tile-count / HF-count / dimension sweep against the oracle, reference quirk Q1 (rows with fidelity not in {0, 1}),
value-only calls, persistent-grid tails, and size-independent properties at the BASELINE batch size."""
import numpy as np
import pytest

from oracle import mfgp_oracle as onp
from oracle import mfgp_oracle_torch as otc
from tests.test_cuda_kernels import rand_theta, rand_X

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h():
    from multi_fidelity_gpflow_b200 import _lib

    return _lib.Handle(0)


def check(h, X, Y, th, nz, rtol_v=1e-9, rtol_g=1e-7):
    nlml, grad = h.gpr_batched_nlml_grad(X, Y, th, nz)
    vals, grads = otc.gpr_batched_value_and_grad(X, Y[:, np.arange(th.shape[0]) % Y.shape[1]], th, nz)  # problem b: column b % ycols
    np.testing.assert_allclose(nlml, -vals, rtol=rtol_v, atol=1e-10)  # NLML within 1e-9 relative (north-star tolerance)
    np.testing.assert_allclose(grad, -grads, rtol=rtol_g, atol=rtol_g * max(1.0, np.abs(grads).max()))  # gradients 1e-7
    return nlml, grad


@pytest.mark.parametrize("N", [7, 8, 9, 16, 24, 31, 40, 48, 53, 56, 57, 64])
@pytest.mark.parametrize("d", [5, 3])  # d = 5 is the statically unrolled instantiation, d = 3 the generic one
def test_every_tile_count_and_both_instantiations(h, N, d):
    rng = np.random.default_rng(100 * N + d)
    X = rand_X(rng, N, d, 0.25)
    B = 7
    Y = rng.standard_normal((N, 3))  # B > ycols: problem b uses column b % 3
    th = np.stack([rand_theta(rng, d) for _ in range(B)])
    check(h, X, Y, th, rng.uniform(1e-3, 1e-1, B))


@pytest.mark.parametrize("nH", [0, 1, 3, 20, 53])
def test_hf_count_sweep_on_hbs_shape(h, nH):
    rng = np.random.default_rng(nH)
    N, d = 53, 5
    X = np.hstack([rng.random((N, d)), np.zeros((N, 1))])
    X[rng.permutation(N)[:nH], d] = 1.0  # HF points anywhere, not only at the end
    Y = rng.standard_normal((N, 4))
    th = np.stack([rand_theta(rng, d) for _ in range(6)])
    check(h, X, Y, th, np.full(6, 1e-3))


def test_dead_rows_quirk_Q1(h):
    """Rows whose fidelity is neither 0.0 nor 1.0 (2.0, 1 + 1 ulp, NaN) have zero covariance (reference linear.py:82):
    their only contribution is the noise on the diagonal."""
    rng = np.random.default_rng(3)
    N, d = 30, 5
    X = rand_X(rng, N, d, 0.3)
    X[4, d], X[11, d], X[29, d] = 2.0, np.nextafter(1.0, 2.0), np.nan
    Y = rng.standard_normal((N, 2))
    th = np.stack([rand_theta(rng, d) for _ in range(4)])
    nz = np.array([1e-3, 1e-2, 0.5, 1e-3])
    nlml, grad = check(h, X, Y, th, nz)
    assert np.all(np.isfinite(nlml)) and np.all(np.isfinite(grad))


def test_rho_equal_to_one_does_not_confuse_fidelity_detection(h):
    rng = np.random.default_rng(4)
    X = rand_X(rng, 53, 5, 0.2)
    Y = rng.standard_normal((53, 2))
    th = np.stack([rand_theta(rng, 5) for _ in range(3)])
    th[:, 0] = 1.0  # rho == 1: the LF row scale equals the HF one
    check(h, X, Y, th, np.full(3, 1e-3))


def test_value_only_and_ragged_persistent_tail(h):
    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"]
    rng = np.random.default_rng(5)
    sms = 148
    for B in (1, 11, 12 * sms - 1, 12 * sms + 5):  # below, at and just above one resident wave of warps
        th = np.tile(onp.default_theta(5), (B, 1)) * np.exp(0.2 * rng.standard_normal((B, 13)))
        nz = np.full(B, 1e-3)
        v_only, g_none = h.gpr_batched_nlml_grad(X, Y, th, nz, want_grad=False)
        assert g_none is None
        v, g = h.gpr_batched_nlml_grad(X, Y, th, nz)
        assert np.array_equal(v, v_only)  # the value does not depend on whether the gradient phases run
        pick = rng.permutation(B)[:6]
        vals, grads = otc.gpr_batched_value_and_grad(X, Y[:, pick % 49], th[pick], nz[pick])
        np.testing.assert_allclose(v[pick], -vals, rtol=1e-9)
        np.testing.assert_allclose(g[pick], -grads, rtol=1e-7, atol=1e-7 * np.abs(grads).max())


def test_full_bench_batch_properties(h):
    """BASELINE config 2 at bench size (49 bins x 16 384 hyper-parameter sets in one launch): properties that need no
    oracle run -- (1) a problem's result does not depend on its position in the batch or on the warp that ran it
    (bit-identical for repeated inputs), (2) sum over bins at theta_b == theta equals the shared-kernel GPR (golden G1),
    (3) spot checks against the oracle."""
    import torch

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"]
    R, nb = 16384, 49
    rng = np.random.default_rng(6)
    base = np.tile(onp.default_theta(5), (nb, 1)) * np.exp(0.3 * rng.standard_normal((nb, 13)))
    th = np.tile(base, (R, 1))  # every restart repeats the same 49 hyper-parameter sets
    th[:nb] = onp.default_theta(5)  # restart 0: the reference's initial values
    nz = np.full(R * nb, 1e-3)
    dev = torch.device("cuda:0")
    tX, tY = torch.from_numpy(X).to(dev), torch.from_numpy(np.ascontiguousarray(Y)).to(dev)
    tth, tnz = torch.from_numpy(th).to(dev), torch.from_numpy(nz).to(dev)
    nlml = torch.empty(R * nb, dtype=torch.float64, device=dev)
    grad = torch.empty(R * nb, 14, dtype=torch.float64, device=dev)
    h.gpr_batched_nlml_grad(tX, tY, tth, tnz, nlml=nlml, grad=grad)
    torch.cuda.synchronize()
    v, g = nlml.cpu().numpy().reshape(R, nb), grad.cpu().numpy().reshape(R, nb, 14)
    assert np.array_equal(v[1:], np.broadcast_to(v[1], (R - 1, nb)))
    assert np.array_equal(g[1:], np.broadcast_to(g[1], (R - 1, nb, 14)))
    g1 = 1825.2620287500806  # golden G1: LML of the HBS data at the initial hyper-parameters
    assert abs(-v[0].sum() - g1) < 1e-9 * g1
    vals, grads = otc.gpr_batched_value_and_grad(X, Y[:, :5], base[:5], nz[:5])
    np.testing.assert_allclose(v[7, :5], -vals, rtol=1e-9)
    np.testing.assert_allclose(g[7, :5], -grads, rtol=1e-7, atol=1e-7 * np.abs(grads).max())


def test_pipelined_host_call_equals_device_call(h):
    """B >= 16 waves with host buffers takes the chunked H2D / compute / D2H pipeline: same bits as the one-launch
    device-buffer call, including the Y column of every chunk's first problem and per-problem info."""
    import torch

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"]
    B = 16 * 148 * 12 + 12345  # ragged last chunk
    rng = np.random.default_rng(9)
    th = np.tile(onp.default_theta(5), (B, 1)) * np.exp(0.25 * rng.standard_normal((B, 13)))
    nz = np.full(B, 1e-3)
    info = np.zeros(B, dtype=np.int32)
    v_h, g_h = h.gpr_batched_nlml_grad(X, Y, th, nz, info=info)
    dev = torch.device("cuda:0")
    tv = torch.empty(B, dtype=torch.float64, device=dev)
    tg = torch.empty(B, 14, dtype=torch.float64, device=dev)
    h.gpr_batched_nlml_grad(torch.from_numpy(X).to(dev), torch.from_numpy(np.ascontiguousarray(Y)).to(dev),
                            torch.from_numpy(th).to(dev), torch.from_numpy(nz).to(dev), nlml=tv, grad=tg)
    torch.cuda.synchronize()
    assert np.array_equal(v_h, tv.cpu().numpy()) and np.array_equal(g_h, tg.cpu().numpy())
    assert not info.any()
    pick = np.array([0, 1, 7103, 7104, B - 1])
    vals, grads = otc.gpr_batched_value_and_grad(X, Y[:, pick % 49], th[pick], nz[pick])
    np.testing.assert_allclose(v_h[pick], -vals, rtol=1e-9)
