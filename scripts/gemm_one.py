"""One rank-K update C = beta*C - A A^T at n x n (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib
n, K, beta, cfg = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else ""
dev = torch.device("cuda:0")
h = _lib.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream); h.set_async(True)
C = torch.zeros(n, n, dtype=torch.float64, device=dev)
A = torch.randn(n, K, dtype=torch.float64, device=dev)
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    rc = _lib._lib.mfgp_gemm(h._h, b"N", b"T", n, n, K, -1.0, _lib._ptr(A), K, _lib._ptr(A), K, beta, _lib._ptr(C), n)
    e1.record(s); torch.cuda.synchronize()
    assert rc == 0
    print(f"rep {rep}: {e0.elapsed_time(e1):.3f} ms {2*n*n*K/(e0.elapsed_time(e1)*1e-3)/1e12:.2f} TF")
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(s); C.mul_(1.0000001); t1.record(s); torch.cuda.synchronize()
print(f"torch RMW of C: {t0.elapsed_time(t1):.3f} ms  {2*C.numel()*8/(t0.elapsed_time(t1)*1e-3)/1e9:.0f} GB/s")
