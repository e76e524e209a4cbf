"""Multi-GPU tests (skipped unless >= 2 CUDA devices): data-parallel SVGP over NCCL equals the single-GPU
result; sharded bins equal the unsharded batch."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from multi_fidelity_gpflow_b200 import _lib
from multi_fidelity_gpflow_b200.dist import dp_svgp_value_and_grad, gather_bins, shard_range
from oracle import mfgp_oracle as onp
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
h = _lib.Handle(rank)
ds = onp.load_dataset("hbs")
X, Y = ds["X"], ds["Y"]
rng = np.random.default_rng(0)
B = Y.shape[1]
th = np.tile(onp.default_theta(5), (B, 1)) * np.exp(0.1 * rng.standard_normal((B, 13)))
nz = np.full(B, 1e-3)
lo, hi = shard_range(B, rank, world)
nl, gr = h.gpr_batched_nlml_grad(X, np.ascontiguousarray(Y[:, lo:hi]), th[lo:hi], nz[lo:hi])
full_n, full_g = gather_bins(nl, B), gather_bins(gr, B)
M, L, P = 50, 10, 49
W = onp.initialize_W(P, L, 0.4, 0.2)
ths = np.tile(onp.default_theta(5), (L, 1))
q_mu, q_sqrt = 0.1 * rng.standard_normal((M, L)), np.tile(0.3 * np.eye(M), (L, 1, 1))
Z = ds["Z_kmeans50"]
fn = lambda Xr, Yr, scale, klm: h.svgp_elbo_grad(Xr, Yr, Z, ths, W, q_mu, q_sqrt, 0.8, scale=scale, kl_mult=klm)
out = dp_svgp_value_and_grad(fn, X, Y, num_data=53, kl_mult=1.3)
# data-parallel TRAINING loop on device memory: two ranks against one rank (sub-group of rank 0) and the plain device loop
import copy
from multi_fidelity_gpflow_b200.kernels import SquaredExponential
from multi_fidelity_gpflow_b200.linear_svgp import LatentMFCoregionalizationSVGP
solo = dist.new_group([0])
mk = lambda: LatentMFCoregionalizationSVGP(X, Y, SquaredExponential(lengthscales=np.ones(5)), SquaredExponential(lengthscales=np.ones(5)),
                                           num_latents=4, num_inducing=20, num_outputs=49, handle=h)
m2 = mk()
m2.optimize_data_parallel((X, Y), max_iters=6, initial_lr=0.01, kl_multiplier=1.3)
checksum = torch.tensor([float(np.sum(m2.q_mu.numpy())), float(m2.loss_history[-1])], dtype=torch.float64, device="cuda")
both = [torch.empty_like(checksum) for _ in range(world)]
dist.all_gather(both, checksum)
if rank == 0:
    m1, m0 = mk(), mk()
    m1.optimize_data_parallel((X, Y), max_iters=6, initial_lr=0.01, kl_multiplier=1.3, group=solo)
    m0.optimize_on_device((X, Y), max_iters=6, initial_lr=0.01, kl_multiplier=1.3)
    dp = {"ranks_identical": bool(torch.equal(both[0], both[1])),
          "loss_2v1": float(np.max(np.abs(np.array(m2.loss_history) / np.array(m1.loss_history) - 1))),
          "loss_1v0": float(np.max(np.abs(np.array(m1.loss_history) / np.array(m0.loss_history) - 1))),
          "qmu_2v1": float(np.max(np.abs(m2.q_mu.numpy() - m1.q_mu.numpy())) / np.max(np.abs(m1.q_mu.numpy())))}
if rank == 0:
    ref_n, ref_g = h.gpr_batched_nlml_grad(X, Y, th, nz)
    ref = h.svgp_elbo_grad(X, Y, Z, ths, W, q_mu, q_sqrt, 0.8, scale=1.0, kl_mult=1.3)
    ok = bool(np.array_equal(full_n, ref_n) and np.array_equal(full_g, ref_g))
    err = {k: float(np.max(np.abs(np.asarray(out[k]) - np.asarray(ref[k]))) / (np.max(np.abs(np.asarray(ref[k]))) + 1e-300))
           for k in ("g_Z", "g_thetas", "g_q_mu", "g_q_sqrt", "g_W", "g_lik_var", "elbo")}
    print("RESULT " + json.dumps({"bins_bit_identical": ok, "err": err, "dp": dp}))
dist.destroy_process_group()
''' % ROOT


def test_two_gpu_sharding_and_dp_svgp(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29631", str(script)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    import json

    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][0]
    r = json.loads(line[7:])
    assert r["bins_bit_identical"]
    assert all(v < 1e-10 for v in r["err"].values()), r
    dp = r["dp"]
    assert dp["ranks_identical"] and dp["loss_2v1"] < 1e-10 and dp["loss_1v0"] < 1e-12 and dp["qmu_2v1"] < 1e-8, dp


def test_two_handles_in_one_process():
    """ADVICE r1 / include/mfgp.h threading contract: one handle per GPU in ONE process.  Function attributes (the > 48 KB
    dynamic shared-memory opt-in of the GEMM, potrf, covariance and K6 kernels) are per device, so the second device must
    get its own opt-in: run every large-smem kernel family on device 1 after device 0 has used them."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from multi_fidelity_gpflow_b200 import _lib
    from oracle import mfgp_oracle as onp

    ds = onp.load_dataset("hbs")
    X, Y = ds["X"], ds["Y"]
    big = onp.synthetic_exact_dataset(1500)
    th = np.tile(onp.default_theta(5), (49, 1))
    res = []
    for dev in (0, 1):
        h = _lib.Handle(dev)
        nl, gr = h.gpr_batched_nlml_grad(X, Y, th, np.full(49, 1e-3))                         # K6
        v, g = h.gpr_nlml_grad(big["X"], big["Y"], big["theta"], big["noise"])               # cov stream, potrf, trtri, gemm, cov_grad
        K = h.cov(big["X"][:200], big["X"][200:500], big["theta"])                           # cov tile kernel
        res.append((nl, gr, v, g, K))
        h.close()
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)  # same kernels, same launch shapes on both devices: bit-identical


CHOL_WORKER = r"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from multi_fidelity_gpflow_b200 import _lib
from multi_fidelity_gpflow_b200.dist_chol import distributed_gpr_nlml
from oracle import mfgp_oracle as onp
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
h = _lib.Handle(rank)
res = {}
for N, nbd in ((1500, 256), (2048, 512)):
    ds = onp.synthetic_exact_dataset(N)
    v = distributed_gpr_nlml(h, ds["X"], ds["Y"], ds["theta"], ds["noise"], nbd=nbd)
    v2, g2 = distributed_gpr_nlml(h, ds["X"], ds["Y"], ds["theta"], ds["noise"], nbd=nbd, want_grad=True)
    if rank == 0:
        h.set_stream(None)
        single = h.gpr_nlml(ds["X"], ds["Y"], ds["theta"], ds["noise"])
        _, gs = h.gpr_nlml_grad(ds["X"], ds["Y"], ds["theta"], ds["noise"])
        ref = -onp.gpr_lml(ds["X"], ds["Y"], ds["theta"], ds["noise"])
        gerr = float(np.max(np.abs(g2 - gs)) / np.max(np.abs(gs)))
        res[N] = [v, single, ref, v2, gerr]
if rank == 0:
    print("RESULT " + json.dumps(res))
dist.destroy_process_group()
""" % ROOT


def test_two_gpu_distributed_cholesky_nlml(tmp_path):
    import json

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "chol_worker.py"
    script.write_text(CHOL_WORKER)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29632", str(script)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    r = json.loads([l for l in res.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    for N, (v, single, ref, v2, gerr) in r.items():
        assert abs(v - ref) < 1e-9 * abs(ref), (N, v, ref)
        assert abs(v - single) < 1e-9 * abs(single)
        assert abs(v2 - v) < 1e-12 * abs(v)  # value identical with and without the gradient phase
        assert gerr < 1e-7, (N, gerr)        # distributed gradient == single-GPU analytic gradient (1e-7 relative)
