"""Exact-GPR predict_f through the C-ABI with host buffers: wall time per call for a few and for many test points."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multi_fidelity_gpflow_b200 import _lib
from oracle import mfgp_oracle as onp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
h = _lib.Handle(0)
ds = onp.synthetic_exact_dataset(N)
rng = np.random.default_rng(0)
for Ns in (100, N // 8, N // 8 + 8, N // 2):
    Xs = ds["X"][rng.integers(0, N, Ns)] + 0.01
    Xs[:, -1] = 1.0
    fn = lambda: h.gpr_predict(ds["X"], ds["Y"], Xs, ds["theta"], ds["noise"])
    fn()
    ts = []
    for _ in range(2):
        t0 = time.perf_counter(); m, v = fn(); ts.append(time.perf_counter() - t0)
    print(f"N={N} Ns={Ns}: {min(ts) * 1e3:.1f} ms  mean[0]={m[0, 0]:.6f} var[0]={v[0]:.3e}")
