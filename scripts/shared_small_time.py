import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from multi_fidelity_gpflow_b200 import _lib
from oracle import mfgp_oracle as onp
h = _lib.Handle(0)
ds = onp.load_dataset("hbs"); X, Y = ds["X"], ds["Y"]; th = onp.default_theta(5)
for _ in range(5): h.gpr_nlml_grad(X, Y, th, 1e-3)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): v, g = h.gpr_nlml_grad(X, Y, th, 1e-3)
torch.cuda.synchronize(); print(f"hbs shared gpr_nlml_grad: {(time.perf_counter()-t0)/200*1e6:.1f} us/eval  nlml={v:.10f}")
