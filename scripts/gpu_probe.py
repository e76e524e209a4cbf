"""On-box probe: FP64 pipe peaks (DFMA / DMMA microbenchmarks), cuBLAS DGEMM and cuSOLVER
potrf comparators (through torch), and this library's GEMM / potrf / cov rates.
Writes gpurun_out/probe.json.  Library comparators are measurement only, never product."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib  # noqa: E402


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    h = _lib.Handle(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    out["dfma_tflops"] = h.fp64_peak(0, 40000) / 1e12
    out["dmma_tflops"] = h.fp64_peak(1, 40000) / 1e12
    print(out, flush=True)
    dev = torch.device("cuda:0")
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        c = torch.empty_like(a)
        t = ev_time(lambda: torch.matmul(a, b, out=c))
        out[f"cublas_dgemm_{n}_tflops"] = 2 * n**3 / t / 1e12
        h.set_async(True)
        for ta, tb in ((False, True), (False, False), (True, False)):
            t = ev_time(lambda: _lib._lib.mfgp_gemm(h._h, b"T" if ta else b"N", b"T" if tb else b"N", n, n, n, 1.0,
                                                    _lib._ptr(a), n, _lib._ptr(b), n, 0.0, _lib._ptr(c), n))
            out[f"mfgp_dgemm_{'T' if ta else 'N'}{'T' if tb else 'N'}_{n}_tflops"] = 2 * n**3 / t / 1e12
        h.set_async(False)
        print(out, flush=True)
        del a, b, c
    for n in (4096, 8192, 16384):
        x = torch.randn(n, 64, dtype=torch.float64, device=dev)
        a = x @ x.T + n * torch.eye(n, dtype=torch.float64, device=dev)
        t = ev_time(lambda: torch.linalg.cholesky(a), reps=2)
        out[f"cusolver_potrf_{n}_tflops"] = n**3 / 3 / t / 1e12
        work = a.clone()
        h.set_async(True)

        def run():
            work.copy_(a)
            h.potrf_device(work, n, n)

        tc = ev_time(lambda: work.copy_(a), reps=2)
        t = ev_time(run, reps=2) - tc
        h.set_async(False)
        assert h.sync() == 0
        out[f"mfgp_potrf_{n}_tflops"] = n**3 / 3 / t / 1e12
        ref = torch.linalg.cholesky(a)
        out[f"mfgp_potrf_{n}_relerr"] = float((torch.tril(work) - ref).abs().max() / ref.abs().max())
        print(out, flush=True)
        del a, work, ref, x
    # covariance assembly bandwidth
    for n, d in ((8192, 10), (16384, 10)):
        X = torch.rand(n, d + 1, dtype=torch.float64, device=dev)
        X[:, -1] = (torch.arange(n, device=dev) >= n * 7 // 8).double()
        th = torch.ones(2 * d + 3, dtype=torch.float64, device=dev)
        K = torch.empty(n, n, dtype=torch.float64, device=dev)
        h.set_async(True)
        t = ev_time(lambda: _lib._lib.mfgp_cov(h._h, _lib._ptr(X), n, None, n, d, _lib._ptr(th), _lib._ptr(K), n))
        out[f"mfgp_cov_sym_{n}_GBs"] = (8.0 * n * n + 8.0 * 2 * n * (d + 1)) / t / 1e9
        X2 = X.clone()
        t = ev_time(lambda: _lib._lib.mfgp_cov(h._h, _lib._ptr(X), n, _lib._ptr(X2), n, d, _lib._ptr(th), _lib._ptr(K), n))
        out[f"mfgp_cov_rect_{n}_GBs"] = (8.0 * n * n + 8.0 * 2 * n * (d + 1)) / t / 1e9
        h.set_async(False)
        del K
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
