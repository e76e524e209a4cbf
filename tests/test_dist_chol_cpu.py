"""CPU (gloo) tests of the 2-D block-cyclic distributed Cholesky SCHEDULE: ownership, panel broadcasts, look-ahead
ordering, ragged N.  The block arithmetic is a torch-CPU/NumPy double defined here (test infrastructure, built on the
oracle); the product path runs the same schedule on libmfgp.so (tests/test_multi_gpu.py)."""
import contextlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class CpuOpsDouble:
    """Same interface as dist_chol.GpuOps, torch CPU tensors, oracle covariance."""

    def __init__(self):
        import torch

        self.torch = torch
        self.device = torch.device("cpu")

    def new_stream(self, high_priority=False):
        return None

    def use(self, stream):
        return contextlib.nullcontext()

    def record(self, stream):
        return None

    def wait(self, stream, event):
        pass

    def finish(self):
        pass

    def cov(self, Xa, Xb, theta, out):
        from oracle import mfgp_oracle as onp

        out.copy_(self.torch.from_numpy(onp.mf_K(Xa.numpy(), Xb.numpy(), theta.numpy())))

    def cov_grad(self, X, theta, G, scale, out):
        """scale * sum_ij Gs_ij dK_ij/dtheta with Gs = symmetric completion of the lower triangle of G; last entry trace."""
        from oracle import mfgp_oracle_torch as otc

        torch = self.torch
        Gs = torch.tril(G) + torch.tril(G, -1).T
        th = theta.clone().requires_grad_(True)
        K = otc.mf_K(X, X, th)
        (Gs * K).sum().backward()
        out[:-1] = scale * th.grad
        out[-1] = scale * torch.trace(Gs)

    def potrf_inv(self, A, W):
        L = self.torch.linalg.cholesky(A)
        A.copy_(L)
        W.copy_(self.torch.linalg.inv(L))

    def gemm(self, ta, tb, m, n, k, alpha, A, B, beta, C):
        a = A[:k, :m].T if ta else A[:m, :k]
        b = B[:n, :k].T if tb else B[:k, :n]
        C[:m, :n] = alpha * (a @ b) + (beta * C[:m, :n] if beta != 0.0 else 0.0)  # beta == 0 must not read C (BLAS rule)


def _worker(rank, world, port, grid, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from multi_fidelity_gpflow_b200.dist_chol import distributed_gpr_nlml
    from oracle import mfgp_oracle as onp

    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {}
    for N, nb, la in ((96, 16, True), (96, 16, False), (150, 32, True), (64, 64, True), (200, 8, True)):
        ds = onp.synthetic_exact_dataset(N, d=3)
        out[(N, nb, la)] = distributed_gpr_nlml(CpuOpsDouble(), ds["X"], ds["Y"], ds["theta"], ds["noise"], nbd=nb, grid=grid,
                                                lookahead=la)
    for N, nb in ((96, 16), (150, 32), (200, 8)):  # value + gradient (ragged N pads the last block; 25 blocks: many owned rows)
        ds = onp.synthetic_exact_dataset(N, d=3)
        v, g = distributed_gpr_nlml(CpuOpsDouble(), ds["X"], ds["Y"], ds["theta"], ds["noise"], nbd=nb, grid=grid, want_grad=True)
        out[("grad", N, nb)] = (v, g.tolist())
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,grid", [(2, (1, 2)), (2, (2, 1)), (4, (2, 2)), (4, None), (6, (2, 3))])
def test_block_cyclic_schedule_matches_oracle(world, grid):
    import torch.multiprocessing as mp

    from oracle import mfgp_oracle as onp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket

    with socket.socket() as sk:  # a port that is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(60)
    from oracle import mfgp_oracle_torch as otc

    for key, v in out.items():
        if key[0] == "grad":
            _, N, nb = key
            ds = onp.synthetic_exact_dataset(N, d=3)
            lml, gth, gnz = otc.gpr_lml_value_and_grad(ds["X"], ds["Y"], ds["theta"], ds["noise"])
            ref = -np.concatenate([gth, [gnz]])
            assert abs(v[0] + lml) < 1e-9 * abs(lml)
            np.testing.assert_allclose(np.array(v[1]), ref, rtol=1e-7, atol=1e-7 * np.abs(ref).max())
            continue
        N, nb, la = key
        ds = onp.synthetic_exact_dataset(N, d=3)
        ref = -onp.gpr_lml(ds["X"], ds["Y"], ds["theta"], ds["noise"])
        assert abs(v - ref) < 1e-9 * abs(ref), (N, nb, la, v, ref)


def test_process_grid_shapes():
    from multi_fidelity_gpflow_b200.dist_chol import process_grid

    assert process_grid(8) == (2, 4) and process_grid(4) == (2, 2) and process_grid(2) == (1, 2) and process_grid(1) == (1, 1)
    assert process_grid(6) == (2, 3)
