"""Exact-GPR NLML (value only) and NLML + gradient through the C-ABI with host buffers: wall time per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multi_fidelity_gpflow_b200 import _lib
from oracle import mfgp_oracle as onp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
h = _lib.Handle(0)
ds = onp.synthetic_exact_dataset(N)
for name, fn in (("nlml", lambda: h.gpr_nlml(ds["X"], ds["Y"], ds["theta"], ds["noise"])),
                 ("nlml+grad", lambda: h.gpr_nlml_grad(ds["X"], ds["Y"], ds["theta"], ds["noise"])[0])):
    fn()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); v = fn(); ts.append(time.perf_counter() - t0)
    print(f"N={N} {name}: {min(ts) * 1e3:.1f} ms  value {v:.9f}")
