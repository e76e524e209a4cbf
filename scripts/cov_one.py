import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib
n, d, sym = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
h = _lib.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream); h.set_async(True)
X = torch.rand(n, d + 1, dtype=torch.float64, device=dev)
X[:, -1] = (torch.arange(n, device=dev) >= n * 7 // 8).double()
X2 = X.clone()
th = torch.ones(2 * d + 3, dtype=torch.float64, device=dev)
K = torch.empty(n, n, dtype=torch.float64, device=dev)
for _ in range(3):
    assert _lib._lib.mfgp_cov(h._h, _lib._ptr(X), n, None if sym else _lib._ptr(X2), n, d, _lib._ptr(th), _lib._ptr(K), n) == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record(s)
for _ in range(reps):
    assert _lib._lib.mfgp_cov(h._h, _lib._ptr(X), n, None if sym else _lib._ptr(X2), n, d, _lib._ptr(th), _lib._ptr(K), n) == 0
e1.record(s)
torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 1e-3 / reps
print(f"cov N={n} d={d} {'sym' if sym else 'rect'}: {t * 1e3:.3f} ms  {8.0 * n * n / t / 1e9:.0f} GB/s")
