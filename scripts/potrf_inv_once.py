"""Time mfgp_potrf_inv (factor + inverse of the factor) on one n x n block: the per-step critical kernel sequence of the
distributed Cholesky."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib
dev = torch.device("cuda:0")
h = _lib.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream); h.set_async(True)
for n in (256, 512, 1024, 2048):
    x = torch.randn(n, 64, dtype=torch.float64, device=dev)
    a = x @ x.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    w = torch.empty_like(a); A = a.clone()
    best = 1e9
    for rep in range(5):
        A.copy_(a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        assert _lib._lib.mfgp_potrf_inv(h._h, _lib._ptr(A), n, n, _lib._ptr(w), n) == 0
        e1.record(s); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    err = float((torch.tril(w) @ torch.tril(A) - torch.eye(n, dtype=torch.float64, device=dev)).abs().max())
    print(f"potrf_inv n={n}: {best*1e3:.0f} us  ({2*n**3/3/best/1e9:.2f} TFLOP/s)  |W L - I| = {err:.1e}", flush=True)
