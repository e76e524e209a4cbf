"""Time skinny products C[M,2] (+)= A[M,K] B[K,2] through mfgp_gemm (forward-substitution shapes of dist_chol)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib
dev = torch.device("cuda:0")
h = _lib.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream); h.set_async(True)
for M, K in ((1024, 512), (1024, 1024), (15360, 512), (15360, 1024), (15360, 2048), (15872, 512)):
    A = torch.randn(M, K, dtype=torch.float64, device=dev)
    B = torch.randn(K, 2, dtype=torch.float64, device=dev)
    C = torch.zeros(M, 2, dtype=torch.float64, device=dev)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        rc = _lib._lib.mfgp_gemm(h._h, b"N", b"N", M, 2, K, -1.0, _lib._ptr(A), K, _lib._ptr(B), 2, 1.0, _lib._ptr(C), 2)
        e1.record(s); torch.cuda.synchronize()
        assert rc == 0
    ref = -(A @ B) * 3
    print(f"M={M} K={K}: {e0.elapsed_time(e1):.3f} ms  maxerr {float((C - ref).abs().max()):.2e}", flush=True)
