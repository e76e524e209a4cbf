"""One Goku per-bin batched NLML+grad (64 bins, N = 1164) and one single-bin SVGP ELBO+grad; for launch lists / timing."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib
from oracle import mfgp_oracle as onp
which = sys.argv[1] if len(sys.argv) > 1 else "gpr"
h = _lib.Handle(0)
ds = onp.load_dataset("goku")
X, Y = ds["X"], ds["Y"]
N, P = Y.shape
th = onp.default_theta(10)
ths, nz = np.tile(th, (P, 1)), np.full(P, 1e-3)
if which == "gpr":
    fn = lambda: h.gpr_batched_nlml_grad(X, Y, ths, nz)
else:
    M = 300; Z = ds["Z_kmeans300"]; q_mu, q_sqrt = np.zeros((M, P)), np.tile(0.1 * np.eye(M), (P, 1, 1))
    fn = lambda: h.svgp_elbo_grad(X, Y, Z, ths, None, q_mu, q_sqrt, 1.0)
for r in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(f"{which} rep {r}: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
