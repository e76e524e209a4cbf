import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib
dev = torch.device("cuda:0")
h = _lib.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream); h.set_async(True)
def bench(n, K, beta, ldc, reps=10):
    C = torch.zeros(n, n, dtype=torch.float64, device=dev)
    A = torch.randn(n, K, dtype=torch.float64, device=dev)
    def run():
        rc = _lib._lib.mfgp_gemm(h._h, b"N", b"T", n, n, K, -1.0, _lib._ptr(A), K, _lib._ptr(A), K, beta, _lib._ptr(C), ldc)
        assert rc == 0
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps): run()
    e1.record(s); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / reps
    print(f"n={n} K={K} beta={beta} ldc={ldc}: {t*1e3:8.3f} ms {2*n*n*K/t/1e12:6.2f} TF", flush=True)
for n in (2048, 4096, 8192):
    for beta in (0.0, 1.0):
        bench(n, 128, beta, n)
bench(8192, 128, 1.0, 0)
bench(8192, 128, 0.0, 0)
