"""Distributed exact-GP NLML for ONE large problem across the GPUs of a box.

Round-1 form: 1-D block-cyclic by block COLUMNS (width `nbd`), right-looking:
    owner(k): assemble-on-the-fly is done up front by every rank for its own block columns from the
              replicated X (no communication); then per step
              potrf + inverse of the diagonal block, panel solve  (mfgp_potrf_inv, mfgp_gemm)
    all     : panel broadcast (NCCL over NVLink, torch.distributed.broadcast)
    all     : trailing update of the block columns they own       (mfgp_gemm, DMMA)
followed by a distributed forward substitution for a = L^-1 y (the owner of block k broadcasts
[a_k ; L[k+1:,k] a_k]).  The 2-D grid with look-ahead is the round-2 step.  All device arithmetic goes
through libmfgp.so on torch CUDA tensors; torch is only buffers + NCCL."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib


def distributed_gpr_nlml(handle, X, Y, theta, noise, nbd=512, group=None):
    """Every rank passes the same host X [N, d+1], Y [N, 1], theta, noise.  Returns the NLML (same on all ranks)."""
    import torch
    import torch.distributed as dist

    L = _lib._lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, d = X.shape[0], X.shape[1] - 1
    assert Y.shape[1] == 1, "single-output problem (SURVEY config C5)"
    nblk = (N + nbd - 1) // nbd
    Xd = torch.from_numpy(X).to(dev)
    thd = torch.from_numpy(np.ascontiguousarray(theta, dtype=np.float64)).to(dev)
    y = torch.zeros(N, 2, dtype=torch.float64, device=dev)  # column 1 is padding (even leading dimension for the GEMM)
    y[:, 0] = torch.from_numpy(np.ascontiguousarray(Y[:, 0], dtype=np.float64)).to(dev)
    ptr = _lib._ptr
    h = handle._h
    stream = torch.cuda.current_stream()
    handle.set_stream(stream.cuda_stream)
    handle.set_async(True)

    def chk(rc, what):
        if rc != 0:
            raise _lib.MFGPError(f"{what}: rc={rc}: {L.mfgp_last_error(h).decode()}")

    # ---- assembly: block column j (rows j*nbd.., width wj) = K(X[rows], X[cols]) + noise on its diagonal ----
    cols = {}
    for j in range(rank, nblk, world):
        r0, wj = j * nbd, min(nbd, N - j * nbd)
        A = torch.empty(N - r0, nbd, dtype=torch.float64, device=dev)
        chk(L.mfgp_cov(h, ptr(Xd[r0:]), N - r0, ptr(Xd[r0:r0 + wj]), wj, d, ptr(thd), ptr(A), nbd), "cov")
        A[:wj, :wj].diagonal().add_(noise)
        cols[j] = A
    logdet = torch.zeros(1, dtype=torch.float64, device=dev)
    panel = torch.empty(N, nbd, dtype=torch.float64, device=dev)
    Wk = torch.empty(nbd, nbd, dtype=torch.float64, device=dev)
    avec = torch.zeros(N, dtype=torch.float64, device=dev)
    msg = torch.zeros(N, 2, dtype=torch.float64, device=dev)
    for k in range(nblk):
        r0, wk = k * nbd, min(nbd, N - k * nbd)
        own = k % world
        rows = N - r0
        if rank == own:
            A = cols[k]
            # diagonal block: L_kk (in place) and W_kk = L_kk^-1
            chk(L.mfgp_potrf_inv(h, ptr(A), wk, nbd, ptr(Wk), nbd), "potrf_inv")
            logdet += torch.log(A[:wk, :wk].diagonal()).sum()
            if rows > wk:  # panel rows below: L_ik = A_ik W_kk^T (in place: one 128-wide column tile per CTA row)
                tmp = torch.empty(rows - wk, nbd, dtype=torch.float64, device=dev)
                chk(L.mfgp_gemm(h, b"N", b"T", rows - wk, wk, wk, 1.0, ptr(A[wk:]), nbd, ptr(Wk), nbd, 0.0, ptr(tmp), nbd), "panel")
                A[wk:, :wk] = tmp[:, :wk]
            panel[:rows].copy_(A)
            # forward substitution piece: a_k = W_kk y_k ; u = L[k+1:, k] a_k
            chk(L.mfgp_gemm(h, b"N", b"N", wk, 2, wk, 1.0, ptr(Wk), nbd, ptr(y[r0:]), 2, 0.0, ptr(msg), 2), "a_k")
            if rows > wk:
                chk(L.mfgp_gemm(h, b"N", b"N", rows - wk, 2, wk, 1.0, ptr(A[wk:]), nbd, ptr(msg), 2, 0.0, ptr(msg[wk:]), 2), "u")
        dist.broadcast(panel[:rows], src=own, group=group)
        dist.broadcast(msg[:rows], src=own, group=group)
        avec[r0:r0 + wk] = msg[:wk, 0]
        if rows > wk:
            y[r0 + wk:, 0] -= msg[wk:rows, 0]
        # trailing update of the owned block columns j > k:  A_j -= P[j rows] P[j block rows]^T
        for j in range(k + 1, nblk):
            if j % world != rank:
                continue
            o = (j - k) * nbd
            wj = min(nbd, N - j * nbd)
            chk(L.mfgp_gemm(h, b"N", b"T", rows - o, wj, wk, -1.0, ptr(panel[o:]), nbd, ptr(panel[o:]), nbd, 1.0,
                            ptr(cols[j]), nbd), "update")
    handle.set_async(False)
    info = handle.sync()
    if info:
        raise _lib.NotPositiveDefiniteError(f"distributed potrf: pivot {info} not positive")
    dist.all_reduce(logdet, group=group)
    quad = float((avec * avec).sum().item())
    return 0.5 * quad + float(logdet.item()) + 0.5 * N * math.log(2.0 * math.pi)
