"""One-process-per-GPU partitioning of the hot path (torch.distributed; NCCL on GPUs, gloo in the
CPU tests).  How each sub-path shards:

* independent bins / restarts  -> contiguous slices per rank, NO data-path collective
  (`shard_range`, `gather_bins` only to assemble the result vector);
* SVGP minibatch ELBO          -> rows of (X, Y) shard across ranks; every rank evaluates its rows
  with the global scale and kl_mult / world, then ONE all-reduce(sum) of the flat
  [loss, kl, gradients] vector (`dp_svgp_value_and_grad`).
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of n items for `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist

    return dist


def gather_bins(local: np.ndarray, n_total: int, group=None) -> np.ndarray:
    """All-gather per-rank slices (along axis 0) produced with shard_range back into the full array."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    counts = [shard_range(n_total, r, world) for r in range(world)]
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    mx = max(hi - lo for lo, hi in counts)
    buf = torch.zeros(mx * width, dtype=torch.float64, device=dev)
    buf[: local.size] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64).ravel()).to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    parts = [o.cpu().numpy()[: (hi - lo) * width].reshape((hi - lo,) + local.shape[1:]) for o, (lo, hi) in zip(outs, counts)]
    return np.concatenate(parts, axis=0)


def dp_svgp_value_and_grad(local_fn, X, Y, num_data, kl_mult=1.0, group=None):
    """Data-parallel SVGP objective.

    local_fn(X_rows, Y_rows, scale, kl_mult) -> dict(elbo, kl, g_* ...) evaluated on this rank's rows
    (mfgp_svgp_elbo_grad through Handle.svgp_elbo_grad).  Returns the same dict for the GLOBAL batch:
    identical (to rounding) on every rank after one all-reduce.
    """
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = X.shape[0]
    lo, hi = shard_range(B, rank, world)
    scale = float(num_data) / B if num_data is not None else 1.0
    r = local_fn(X[lo:hi], Y[lo:hi], scale, kl_mult / world)
    keys = [k for k in sorted(r) if k.startswith("g_") and r[k] is not None]
    # elbo_r = scale*VE_r - KL  ->  sum_r (elbo_r + KL) - KL
    flat = np.concatenate([[r["elbo"] + r["kl"]]] + [np.ravel(np.asarray(r[k], dtype=np.float64)) for k in keys])
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(flat).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    flat = t.cpu().numpy()
    out = {"kl": r["kl"], "elbo": float(flat[0]) - r["kl"]}
    o = 1
    for k in keys:
        n = int(np.size(r[k]))
        out[k] = flat[o:o + n].reshape(np.shape(r[k])) if np.ndim(r[k]) else float(flat[o])
        o += n
    for k in r:
        if k.startswith("g_") and r[k] is None:
            out[k] = None
    return out
